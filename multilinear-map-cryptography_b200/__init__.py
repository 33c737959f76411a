"""B200 (sm_100a) prover backend for the Twist & Shout hot path.

The directory name carries a hyphen, so import it with
    importlib.import_module("multilinear-map-cryptography_b200")
(tests/conftest.py and bench.py do exactly that).  Contents:
    csrc/       CUDA kernels + the C ABI (include/tsgpu.h)  -> libtsgpu.so
    host/       C++ host mirror of the reference API (Transcript, SumCheck, KZG, Twist, Shout)
    binding.py  ctypes binding of libtsgpu.so - no CPU fallback
"""
from .binding import (Context, Table, SumCheckRounds, SumCheck, SumCheckProof, Transcript, Srs, Poly,  # noqa: F401
                      KZGCommitment, g1_hash, g1_compress, g1_equal, unique_id, chacha20_u64, chacha20_fr_then_u64, statement_digest, TwistAndShoutError, LIB_PATH, lib)
from .api import (setup_params, ProverParams, VerifierParams, MemoryTrace, MemoryOp, Twist, TwistProof,  # noqa: F401,E402
                  LookupTable, LookupOp, Shout, ShoutProof, fe, fe_vec, fe_to_int, HostVerifierParams, kzg_verify, kzg_batch_verify, KZGVectorCommitment,
                  MultilinearExtension, LessThanPolynomial, ShoutReadCheck, TwistMemoryCheck,
                  fe_from_int, fe_add, fe_mul, fe_inverse, field_utils, poly_utils, polynomial_division)
