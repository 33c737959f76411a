"""Python mirror of the reference crate's public API for the prover path (src/lib.rs:50-56 re-exports):
setup_params, MemoryTrace / Twist, LookupTable / Shout - same names, argument meaning and error behaviour,
so the parity tests read like the reference's own tests.  All heavy work happens in libtsgpu.so.

FieldElement values are numpy uint64[4] Montgomery limbs (ark_bn254::Fr's in-memory form); `fe(x)` is
`FieldElement::from(x as u64)`."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np

from .binding import Context, Srs, SumCheckProof, Table, Transcript, TwistAndShoutError, _fr, _p, lib


def fe(x: int) -> np.ndarray:
    """FieldElement::from(x as u64)"""
    v = np.array([x], dtype=np.uint64)
    out = np.empty((1, 4), dtype=np.uint64)
    lib().tsgpu_fr_from_u64(_p(v), C.c_size_t(1), _p(out))
    return out[0]


def fe_vec(xs) -> np.ndarray:
    v = np.ascontiguousarray(xs, dtype=np.uint64).reshape(-1)
    out = np.empty((v.shape[0], 4), dtype=np.uint64)
    lib().tsgpu_fr_from_u64(_p(v), C.c_size_t(v.shape[0]), _p(out))
    return out


def fe_to_int(x) -> int:
    x = _fr(x, 1)
    out = np.empty_like(x)
    lib().tsgpu_fr_to_canonical(_p(x), C.c_size_t(1), _p(out))
    return sum(int(out[0, i]) << (64 * i) for i in range(4))


class MultilinearExtension:
    """src/polynomials.rs:18-196 - the dense table lives in HBM (a `Table`); `num_vars` and `evaluations` read like the reference's
    public fields.  Where the reference panics (asserts) this mirror raises: ValueError for the constructor's length assert,
    TwistAndShoutError(Polynomial) with the reference's message for the others."""

    def __init__(self, table: Table):
        self.table = table

    @property
    def ctx(self) -> Context:
        return self.table.ctx

    @property
    def num_vars(self) -> int:
        return self.table.num_vars

    @property
    def evaluations(self) -> np.ndarray:
        return self.table.download()

    @staticmethod
    def from_evaluations(ctx: Context, evaluations) -> "MultilinearExtension":
        return MultilinearExtension(ctx.table_upload(evaluations))                       # polynomials.rs:28-37 (length must be a power of two)

    @staticmethod
    def from_evaluations_vec(ctx: Context, num_vars: int, evaluations) -> "MultilinearExtension":
        return MultilinearExtension(ctx.table_upload(evaluations, num_vars))             # polynomials.rs:40-50 (pad / truncate)

    @staticmethod
    def from_sparse(ctx: Context, num_vars: int, sparse_entries) -> "MultilinearExtension":
        return MultilinearExtension(ctx.table_from_sparse(num_vars, sparse_entries))     # polynomials.rs:52-67

    @staticmethod
    def one_hot(ctx: Context, num_vars: int, index: int) -> "MultilinearExtension":
        return MultilinearExtension(ctx.table_one_hot(num_vars, index))                  # polynomials.rs:71-82

    def evaluate(self, point) -> np.ndarray:
        point = np.ascontiguousarray(point, dtype=np.uint64).reshape(-1, 4)
        if point.shape[0] != self.num_vars:
            raise TwistAndShoutError(5, "Point dimension must match number of variables")   # polynomials.rs:86
        return self.table.evaluate(point)

    def partial_evaluate(self, fixed_vars) -> "MultilinearExtension":
        return MultilinearExtension(self.table.partial_evaluate(fixed_vars))             # polynomials.rs:126-161

    def add(self, other: "MultilinearExtension") -> "MultilinearExtension":
        return MultilinearExtension(self.table.add(other.table))                         # polynomials.rs:164-176

    def scalar_mul(self, scalar) -> "MultilinearExtension":
        return MultilinearExtension(self.table.scalar_mul(scalar))                       # polynomials.rs:179-189

    def sum_evaluations(self) -> np.ndarray:
        return self.table.sum_evaluations()                                              # polynomials.rs:192-195


class LessThanPolynomial:
    """src/polynomials.rs:201-293"""

    def __init__(self, num_vars: int):
        self.num_vars = num_vars

    @staticmethod
    def new(num_vars: int) -> "LessThanPolynomial":
        return LessThanPolynomial(num_vars)

    def evaluate_at_bits(self, a_bits, b_bits) -> np.ndarray:
        if len(a_bits) != self.num_vars or len(b_bits) != self.num_vars:
            raise ValueError("bit vectors must have num_vars entries")                    # assert_eq!, polynomials.rs:224-225
        for x, y in zip(a_bits, b_bits):                                                  # polynomials.rs:229-238
            if x and not y:
                return fe(0)
            if (not x) and y:
                return fe(1)
        return fe(0)

    def evaluate_at_field_elements(self, a, b) -> np.ndarray:
        a = _fr(a, 1); b = _fr(b, 1)
        out = np.empty(4, dtype=np.uint64)
        lib().tsgpu_lt_evaluate_at_field_elements(C.c_uint(self.num_vars), _p(a), _p(b), _p(out))
        return out

    def to_multilinear_extension(self, ctx: Context) -> MultilinearExtension:
        return MultilinearExtension(ctx.table_less_than(self.num_vars))                  # polynomials.rs:243-263


class ProverParams:
    """src/utils.rs:21-34 (+ the device-resident SRS).  VerifierParams shares the object: the fields the
    reference copies into it (log_size, max_operations, fiat_shamir_seed) are the same values."""

    def __init__(self, ctx: Context, handle: C.c_void_p):
        self.ctx = ctx
        self._h = handle
        L = lib()
        L.tsgpu_params_log_size.restype = C.c_size_t; L.tsgpu_params_log_size.argtypes = [C.c_void_p]
        L.tsgpu_params_max_operations.restype = C.c_size_t; L.tsgpu_params_max_operations.argtypes = [C.c_void_p]
        L.tsgpu_params_srs.restype = C.c_void_p; L.tsgpu_params_srs.argtypes = [C.c_void_p]
        L.tsgpu_params_tau.argtypes = [C.c_void_p, C.c_void_p]
        L.tsgpu_params_fiat_shamir_seed.argtypes = [C.c_void_p, C.c_void_p]
        L.tsgpu_params_free.argtypes = [C.c_void_p, C.c_void_p]
        self.log_size = int(L.tsgpu_params_log_size(handle))
        self.max_operations = int(L.tsgpu_params_max_operations(handle))
        tau = np.empty(4, dtype=np.uint64); L.tsgpu_params_tau(handle, _p(tau))
        self.tau = tau
        seed = np.empty(32, dtype=np.uint8); L.tsgpu_params_fiat_shamir_seed(handle, _p(seed))
        self.fiat_shamir_seed = seed.tobytes()

    @property
    def srs(self) -> Srs:
        """commitment_params.g1_powers on the device (borrowed handle: do not free)"""
        s = Srs.__new__(Srs)
        s.ctx = self.ctx
        s._h = C.c_void_p(lib().tsgpu_params_srs(self._h))
        s.free = lambda: None
        return s

    def free(self):
        if self._h:
            lib().tsgpu_params_free(self.ctx._h, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass


VerifierParams = ProverParams


class HostVerifierParams:
    """VerifierParams built without a GPU (src/utils.rs:36-50): what a verifier-only process holds"""

    def __init__(self, log_size: int):
        self._h = C.c_void_p()
        rc = lib().tsgpu_setup_verifier_params(C.c_size_t(log_size), C.byref(self._h))
        if rc:
            raise TwistAndShoutError(rc, "setup_verifier_params failed")
        self.log_size = log_size
        self.max_operations = 4 << log_size

    def __del__(self):
        try:
            if self._h:
                lib().tsgpu_params_free(None, self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


def kzg_verify(verifier_params, commitment, point, value, proof) -> bool:
    """KZGCommitment::verify(vk, commitment, point, value, proof) (src/commitments.rs:201-228); CPU pairing"""
    ok = C.c_int(0)
    commitment = np.ascontiguousarray(commitment, dtype=np.uint64).reshape(12)
    proof = np.ascontiguousarray(proof, dtype=np.uint64).reshape(12)
    rc = lib().tsgpu_kzg_verify(verifier_params._h, _p(commitment), _p(_fr(point, 1)), _p(_fr(value, 1)), _p(proof), C.byref(ok))
    if rc:
        raise TwistAndShoutError(rc, "kzg_verify: bad arguments")
    return bool(ok.value)


class KZGVectorCommitment:
    """Mirror of `impl VectorCommitmentScheme for KZGVectorCommitment` (src/commitments.rs:407-483): associated functions.
    `params` is an Srs handle; vectors are host arrays uint64[n, 4] of any length."""

    @staticmethod
    def commit(params: Srs, vector) -> np.ndarray:
        ctx = params.ctx
        vector = _fr(vector) if len(vector) else np.empty((0, 4), dtype=np.uint64)
        out = np.empty(12, dtype=np.uint64)
        ctx.check(lib().tsgpu_vector_commit(ctx._h, params._h, _p(vector), C.c_size_t(vector.shape[0]), _p(out)))
        return out

    @staticmethod
    def open(params: Srs, vector, index: int):
        """-> (value, proof); Commitment("Index out of bounds") beyond the vector"""
        ctx = params.ctx
        vector = _fr(vector) if len(vector) else np.empty((0, 4), dtype=np.uint64)
        value = np.empty(4, dtype=np.uint64); proof = np.empty(12, dtype=np.uint64)
        ctx.check(lib().tsgpu_vector_open(ctx._h, params._h, _p(vector), C.c_size_t(vector.shape[0]), C.c_size_t(index), _p(value), _p(proof)))
        return value, proof

    @staticmethod
    def verify(verifier_params, commitment, index: int, value, proof) -> bool:
        ok = C.c_int(0)
        commitment = np.ascontiguousarray(commitment, dtype=np.uint64).reshape(12)
        proof = np.ascontiguousarray(proof, dtype=np.uint64).reshape(12)
        rc = lib().tsgpu_vector_verify(verifier_params._h, _p(commitment), C.c_size_t(index), _p(_fr(value, 1)), _p(proof), C.byref(ok))
        if rc:
            raise TwistAndShoutError(rc, "vector verify: bad arguments")
        return bool(ok.value)


def kzg_batch_verify(verifier_params, commitments, points, values, proofs) -> bool:
    """KZGCommitment::batch_verify (src/commitments.rs:230-301)"""
    commitments = np.ascontiguousarray(commitments, dtype=np.uint64).reshape(-1, 12)
    proofs = np.ascontiguousarray(proofs, dtype=np.uint64).reshape(-1, 12)
    points = _fr(points); values = _fr(values)
    n = commitments.shape[0]
    if not (points.shape[0] == values.shape[0] == proofs.shape[0] == n):
        raise TwistAndShoutError(4, "Batch verify input lengths must match")            # commitments.rs:237-243
    ok = C.c_int(0)
    rc = lib().tsgpu_kzg_batch_verify(verifier_params._h, _p(commitments), _p(points), _p(values), _p(proofs), C.c_size_t(n), C.byref(ok))
    if rc:
        raise TwistAndShoutError(rc, "kzg_batch_verify: bad arguments")
    return bool(ok.value)


def setup_params(ctx: Context, log_size: int) -> Tuple[ProverParams, VerifierParams]:
    """setup_params(log_size) -> (ProverParams, VerifierParams)   (src/utils.rs:79-131)"""
    h = C.c_void_p()
    ctx.check(lib().tsgpu_setup_params(ctx._h, C.c_size_t(log_size), C.byref(h)))
    p = ProverParams(ctx, h)
    return p, p


class MemoryOp:
    """src/twist.rs:16-20"""
    __slots__ = ("kind", "address", "value")

    def __init__(self, kind: str, address: int, value: np.ndarray):
        self.kind, self.address, self.value = kind, address, value


class MemoryTrace:
    """src/twist.rs:24-72"""

    def __init__(self, memory_size: int):
        assert memory_size > 0 and memory_size & (memory_size - 1) == 0, "Memory size must be power of 2"   # twist.rs:38
        self.memory_size = memory_size
        self.operations: List[MemoryOp] = []
        self._memory = {}

    @staticmethod
    def new(memory_size: int) -> "MemoryTrace":
        return MemoryTrace(memory_size)

    def write(self, address: int, value) -> None:
        if address >= self.memory_size:
            raise TwistAndShoutError(1, "Address out of bounds")                      # twist.rs:49-53
        value = _fr(value, 1).reshape(4).copy()
        self._memory[address] = value
        self.operations.append(MemoryOp("W", address, value))

    def read(self, address: int) -> np.ndarray:
        if address >= self.memory_size:
            raise TwistAndShoutError(1, "Address out of bounds")                      # twist.rs:62-66
        value = self._memory.get(address)
        if value is None:
            value = np.zeros(4, dtype=np.uint64)
        self.operations.append(MemoryOp("R", address, value))
        return value

    def arrays(self):
        n = len(self.operations)
        addr = np.fromiter((op.address for op in self.operations), dtype=np.uint64, count=n)
        vals = np.stack([op.value for op in self.operations]) if n else np.empty((0, 4), dtype=np.uint64)
        isw = np.fromiter((op.kind == "W" for op in self.operations), dtype=np.uint8, count=n)
        return addr, np.ascontiguousarray(vals, dtype=np.uint64), isw


class Proof:
    """TwistProof / ShoutProof (src/twist.rs:76-89, src/shout.rs:64-79)"""

    def __init__(self, handle: C.c_void_p):
        self._h = handle
        L = lib()
        for name in ("tsgpu_proof_num_rounds", "tsgpu_proof_num_openings"):
            getattr(L, name).restype = C.c_size_t; getattr(L, name).argtypes = [C.c_void_p]
        L.tsgpu_proof_bytes.restype = C.c_size_t; L.tsgpu_proof_bytes.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.tsgpu_proof_commitment.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.tsgpu_proof_round_polynomials.argtypes = [C.c_void_p, C.c_void_p]
        L.tsgpu_proof_final_evaluation.argtypes = [C.c_void_p, C.c_void_p]
        L.tsgpu_proof_opening.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.tsgpu_proof_opening_point.argtypes = [C.c_void_p, C.c_void_p]
        L.tsgpu_proof_set_final_evaluation.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.tsgpu_proof_free.argtypes = [C.c_void_p]

    @property
    def commitments(self) -> np.ndarray:
        out = np.empty((2, 12), dtype=np.uint64)
        for i in range(2):
            lib().tsgpu_proof_commitment(self._h, i, _p(out[i]))
        return out

    @property
    def round_polynomials(self) -> np.ndarray:
        n = int(lib().tsgpu_proof_num_rounds(self._h))
        out = np.empty((max(n, 1), 4, 4), dtype=np.uint64)
        lib().tsgpu_proof_round_polynomials(self._h, _p(out))
        return out[:n]

    @property
    def final_evaluation(self) -> np.ndarray:
        out = np.empty(4, dtype=np.uint64); lib().tsgpu_proof_final_evaluation(self._h, _p(out)); return out

    @property
    def opening_proofs(self) -> np.ndarray:
        n = int(lib().tsgpu_proof_num_openings(self._h))
        out = np.empty((n, 12), dtype=np.uint64); v = np.empty(4, dtype=np.uint64)
        for i in range(n):
            lib().tsgpu_proof_opening(self._h, i, _p(out[i]), _p(v))
        return out

    @property
    def final_evaluations(self) -> np.ndarray:
        n = int(lib().tsgpu_proof_num_openings(self._h))
        out = np.empty((n, 4), dtype=np.uint64); g = np.empty(12, dtype=np.uint64)
        for i in range(n):
            lib().tsgpu_proof_opening(self._h, i, _p(g), _p(out[i]))
        return out

    @property
    def opening_point(self) -> np.ndarray:
        out = np.empty(4, dtype=np.uint64); lib().tsgpu_proof_opening_point(self._h, _p(out)); return out

    def to_bytes(self) -> bytes:
        n = int(lib().tsgpu_proof_bytes(self._h, None, 0))
        buf = np.empty(n, dtype=np.uint8)
        lib().tsgpu_proof_bytes(self._h, _p(buf), n)
        return buf.tobytes()

    def tamper_final_evaluation(self, i: int, value):
        lib().tsgpu_proof_set_final_evaluation(self._h, i, _p(_fr(value, 1)))

    def __del__(self):
        try:
            if self._h:
                lib().tsgpu_proof_free(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


TwistProof = Proof
ShoutProof = Proof


class Twist:
    """src/twist.rs:93-316"""

    def __init__(self, prover_params: ProverParams):
        self.prover_params = prover_params   # the reference clones the params (twist.rs:100-104); a handle is shared here

    @staticmethod
    def new(prover_params: ProverParams) -> "Twist":
        return Twist(prover_params)

    def prove(self, trace: MemoryTrace) -> TwistProof:
        addr, vals, isw = trace.arrays()
        return self.prove_arrays(addr, vals, isw)

    def prove_arrays(self, addresses: np.ndarray, values: np.ndarray, is_write: Optional[np.ndarray] = None) -> TwistProof:
        """same as prove() on a trace given as flat arrays (what the C ABI takes)"""
        ctx = self.prover_params.ctx
        addresses = np.ascontiguousarray(addresses, dtype=np.uint64).reshape(-1)
        values = _fr(values)
        n = addresses.shape[0]
        if is_write is None:
            is_write = np.zeros(n, dtype=np.uint8)
        is_write = np.ascontiguousarray(is_write, dtype=np.uint8)
        h = C.c_void_p()
        ctx.check(lib().tsgpu_twist_prove(ctx._h, self.prover_params._h, _p(addresses), _p(values), _p(is_write), C.c_size_t(n), C.byref(h)))
        return Proof(h)

    def prove_sharded(self, local_addresses: np.ndarray, local_values: np.ndarray, total_operations: int) -> TwistProof:
        """ONE proof sharded over the ranks of the context's communicator (ctx.comm_init*): this rank passes the operations that fall in
        its range [rank m / G, (rank + 1) m / G) of the padded trace (m = next power of two of total_operations).  Every rank gets the
        same proof, byte-identical to prove() on one GPU."""
        ctx = self.prover_params.ctx
        local_addresses = np.ascontiguousarray(local_addresses, dtype=np.uint64).reshape(-1)
        local_values = _fr(local_values) if len(local_values) else np.empty((0, 4), dtype=np.uint64)
        h = C.c_void_p()
        ctx.check(lib().tsgpu_twist_prove_sharded(ctx._h, self.prover_params._h, _p(local_addresses), _p(local_values),
                                                  C.c_size_t(local_addresses.shape[0]), C.c_size_t(total_operations), C.byref(h)))
        return Proof(h)

    def prove_sharded_device(self, local_padded_addresses, local_padded_values, padded_operations: int) -> TwistProof:
        """prove_sharded from this rank's zero-padded slices already resident in HBM (Poly handles of padded_operations / ranks entries; not consumed)"""
        ctx = self.prover_params.ctx
        h = C.c_void_p()
        ctx.check(lib().tsgpu_twist_prove_sharded_dev(ctx._h, self.prover_params._h, local_padded_addresses._h, local_padded_values._h,
                                                      C.c_size_t(padded_operations), C.byref(h)))
        return Proof(h)

    @staticmethod
    def shard_range(total_operations: int, rank: int, world: int):
        """[lo, hi) of the operations rank `rank` passes to prove_sharded"""
        m = 1
        while m < total_operations:
            m <<= 1
        count = m // world
        lo = min(rank * count, total_operations)
        return lo, min(lo + count, total_operations) if rank * count < total_operations else lo

    def prove_device(self, padded_addresses, padded_values) -> TwistProof:
        """prove from two zero-padded value vectors already resident in HBM (Poly handles; consumed)"""
        ctx = self.prover_params.ctx
        h = C.c_void_p()
        ctx.check(lib().tsgpu_twist_prove_dev(ctx._h, self.prover_params._h, padded_addresses._h, padded_values._h, C.byref(h)))
        return Proof(h)

    def verify(self, proof: TwistProof, verifier_params: VerifierParams) -> bool:
        ctx = verifier_params.ctx
        ok = C.c_int(0)
        ctx.check(lib().tsgpu_twist_verify(ctx._h, verifier_params._h, proof._h, C.byref(ok)))
        return bool(ok.value)


class LookupOp:
    """src/shout.rs:17-22"""
    __slots__ = ("index", "value")

    def __init__(self, index: int, value: np.ndarray):
        self.index, self.value = index, value


class LookupTable:
    """src/shout.rs:26-60"""

    def __init__(self, entries):
        self.entries = _fr(entries).copy() if len(entries) else np.empty((0, 4), dtype=np.uint64)
        self.lookups: List[LookupOp] = []

    @staticmethod
    def new(entries) -> "LookupTable":
        return LookupTable(entries)

    def lookup(self, index: int) -> np.ndarray:
        if index >= self.entries.shape[0]:
            raise TwistAndShoutError(1, "Lookup index out of bounds")                 # shout.rs:44-48
        value = self.entries[index]
        self.lookups.append(LookupOp(index, value))
        return value

    def size(self) -> int:
        return self.entries.shape[0]


class Shout:
    """src/shout.rs:83-286"""

    def __init__(self, prover_params: ProverParams):
        self.prover_params = prover_params

    @staticmethod
    def new(prover_params: ProverParams) -> "Shout":
        return Shout(prover_params)

    def prove(self, table: LookupTable) -> ShoutProof:
        idx = np.fromiter((l.index for l in table.lookups), dtype=np.uint64, count=len(table.lookups))
        return self.prove_arrays(table.entries, idx)

    def prove_arrays(self, entries: np.ndarray, lookup_indices: np.ndarray) -> ShoutProof:
        ctx = self.prover_params.ctx
        entries = np.ascontiguousarray(entries, dtype=np.uint64).reshape(-1, 4)
        lookup_indices = np.ascontiguousarray(lookup_indices, dtype=np.uint64).reshape(-1)
        h = C.c_void_p()
        ctx.check(lib().tsgpu_shout_prove(ctx._h, self.prover_params._h, _p(entries), C.c_size_t(entries.shape[0]),
                                          _p(lookup_indices), C.c_size_t(lookup_indices.shape[0]), C.byref(h)))
        return Proof(h)

    def prove_sharded(self, local_entries: np.ndarray, total_entries: int, local_lookup_indices: np.ndarray, total_lookups: int) -> ShoutProof:
        """ONE proof sharded over the ranks of the context's communicator: this rank passes the table entries and the lookup indices that
        fall in its ranges of the padded table / padded lookup vector (Twist.shard_range gives both).  Every rank gets the same proof,
        byte-identical to prove() on one GPU."""
        ctx = self.prover_params.ctx
        local_entries = _fr(local_entries) if len(local_entries) else np.empty((0, 4), dtype=np.uint64)
        local_lookup_indices = np.ascontiguousarray(local_lookup_indices, dtype=np.uint64).reshape(-1)
        h = C.c_void_p()
        ctx.check(lib().tsgpu_shout_prove_sharded(ctx._h, self.prover_params._h, _p(local_entries), C.c_size_t(local_entries.shape[0]),
                                                  C.c_size_t(total_entries), _p(local_lookup_indices), C.c_size_t(local_lookup_indices.shape[0]),
                                                  C.c_size_t(total_lookups), C.byref(h)))
        return Proof(h)

    shard_range = staticmethod(Twist.shard_range)

    def verify(self, proof: ShoutProof, verifier_params: VerifierParams) -> bool:
        ctx = verifier_params.ctx
        ok = C.c_int(0)
        ctx.check(lib().tsgpu_shout_verify(ctx._h, verifier_params._h, proof._h, C.byref(ok)))
        return bool(ok.value)


class ShoutReadCheck:
    """The lookup-correctness sum-check the reference leaves as a stub (src/shout.rs:157-184) - core Shout read-checking,
    rv~(r) = sum_x ra~(x, r) Val~(x).  NOT part of the reference's proofs (non-parity extension): Shout.prove stays
    byte-identical to the reference; this proves that every LookupOp { index, value } of a LookupTable returns entries[index]."""

    def __init__(self, ctx: Context):
        self.ctx = ctx

    @staticmethod
    def _statement(table: LookupTable):
        idx = np.fromiter((l.index for l in table.lookups), dtype=np.uint64, count=len(table.lookups))
        vals = np.stack([np.asarray(l.value, dtype=np.uint64).reshape(4) for l in table.lookups]) if table.lookups else np.empty((0, 4), dtype=np.uint64)
        return idx, vals

    def prove(self, table: LookupTable, transcript: Transcript):
        idx, vals = self._statement(table)
        return self.prove_arrays(table.entries, idx, vals, transcript)

    def prove_arrays(self, entries, lookup_indices, lookup_values, transcript: Transcript):
        """-> (claimed_sum, SumCheckProof, challenges)"""
        entries = _fr(entries) if len(entries) else np.empty((0, 4), dtype=np.uint64)
        idx = np.ascontiguousarray(lookup_indices, dtype=np.uint64).reshape(-1)
        vals = _fr(lookup_values) if len(lookup_values) else np.empty((0, 4), dtype=np.uint64)
        k = max(entries.shape[0] - 1, 0).bit_length()
        claimed = np.zeros(4, dtype=np.uint64); fe = np.zeros(4, dtype=np.uint64)
        rp = np.zeros((max(k, 1), 4, 4), dtype=np.uint64); ch = np.zeros((max(k, 1), 4), dtype=np.uint64)
        self.ctx.check(lib().tsgpu_shout_read_check_prove(self.ctx._h, _p(entries), C.c_size_t(entries.shape[0]), _p(idx), _p(vals), C.c_size_t(idx.shape[0]),
                                                          transcript._h, _p(claimed), _p(rp), _p(fe), _p(ch)))
        return claimed, SumCheckProof(rp[:k], fe), ch[:k]

    def verify(self, table: LookupTable, proof: SumCheckProof, transcript: Transcript) -> bool:
        idx, vals = self._statement(table)
        return self.verify_arrays(table.entries, idx, vals, proof, transcript)

    def verify_arrays(self, entries, lookup_indices, lookup_values, proof: SumCheckProof, transcript: Transcript) -> bool:
        entries = _fr(entries) if len(entries) else np.empty((0, 4), dtype=np.uint64)
        idx = np.ascontiguousarray(lookup_indices, dtype=np.uint64).reshape(-1)
        vals = _fr(lookup_values) if len(lookup_values) else np.empty((0, 4), dtype=np.uint64)
        rp = np.ascontiguousarray(proof.round_polynomials, dtype=np.uint64).reshape(-1, 4, 4)
        fe = _fr(proof.final_evaluation, 1)
        ok = C.c_int(0)
        self.ctx.check(lib().tsgpu_shout_read_check_verify(self.ctx._h, _p(entries), C.c_size_t(entries.shape[0]), _p(idx), _p(vals), C.c_size_t(idx.shape[0]),
                                                           transcript._h, _p(rp), C.c_size_t(rp.shape[0]), _p(fe), C.byref(ok)))
        return bool(ok.value)

    # ---- binding to the KZG commitments of a Shout proof (tsgpu_transcript_bind_proof / tsgpu_shout_commitments_match)
    @staticmethod
    def bind(transcript: Transcript, proof: "ShoutProof") -> Transcript:
        """absorb the proof's table / index commitment hashes (labels of src/shout.rs:129-133) before the sum-check draws anything; returns the transcript"""
        rc = lib().tsgpu_transcript_bind_proof(transcript._h, proof._h, C.c_int(1))
        assert rc == 0, rc
        return transcript

    def commitments_match(self, params, proof: "ShoutProof", table: LookupTable) -> bool:
        """do the proof's two commitments commit to THIS table and THESE lookup indices?  (recomputed on the device from the clear statement)"""
        idx, _ = self._statement(table)
        entries = _fr(table.entries) if len(table.entries) else np.empty((0, 4), dtype=np.uint64)
        ok = C.c_int(0)
        self.ctx.check(lib().tsgpu_shout_commitments_match(self.ctx._h, params._h, proof._h, _p(entries), C.c_size_t(entries.shape[0]), _p(idx),
                                                           C.c_size_t(idx.shape[0]), C.byref(ok)))
        return bool(ok.value)


class TwistMemoryCheck:
    """The memory-consistency sum-checks the reference leaves as a stub (src/twist.rs:181-214): read-checking over (cell, cycle) with its
    Val-evaluation sum-check, then write-checking (Inc consistent with the written values and Val) with its own, all on one transcript.  NOT part of the reference's proofs (non-parity extension): Twist.prove stays
    byte-identical to the reference; this proves that every Read of a MemoryTrace returns the value last written to its address."""

    class Proof:
        def __init__(self, claims, part1: SumCheckProof, part2: SumCheckProof, write_claims=None, part3: Optional[SumCheckProof] = None,
                     part4: Optional[SumCheckProof] = None):
            self.claims, self.read_check, self.val_evaluation = claims, part1, part2
            # write-checking (the third sum-check of Twist) and the Val-evaluation of the claim it ends in
            self.write_claims, self.write_check, self.write_val_evaluation = write_claims, part3, part4

    def __init__(self, ctx: Context):
        self.ctx = ctx

    def prove(self, trace: "MemoryTrace", transcript: Transcript) -> "TwistMemoryCheck.Proof":
        addr, vals, isw = trace.arrays()
        return self.prove_arrays(addr, vals, isw, trace.memory_size, transcript)

    @staticmethod
    def _shape(n: int, memory_size: int):
        k = max(memory_size - 1, 0).bit_length()
        t = max(n - 1, 0).bit_length()
        return k, t

    def prove_arrays(self, addresses, values, is_write, memory_size: int, transcript: Transcript) -> "TwistMemoryCheck.Proof":
        addr = np.ascontiguousarray(addresses, dtype=np.uint64).reshape(-1)
        vals = _fr(values) if len(values) else np.empty((0, 4), dtype=np.uint64)
        isw = np.ascontiguousarray(is_write, dtype=np.uint8).reshape(-1)
        k, t = self._shape(addr.shape[0], memory_size)
        claims = np.zeros((2, 4), dtype=np.uint64)
        r1 = np.zeros((max(k + t, 1), 4, 4), dtype=np.uint64); f1 = np.zeros(4, dtype=np.uint64)
        r2 = np.zeros((max(t, 1), 4, 4), dtype=np.uint64); f2 = np.zeros(4, dtype=np.uint64)
        self.ctx.check(lib().tsgpu_twist_memory_check_prove(self.ctx._h, _p(addr), _p(vals), _p(isw), C.c_size_t(addr.shape[0]), C.c_size_t(memory_size),
                                                            transcript._h, _p(claims), _p(r1), _p(f1), _p(r2), _p(f2)))
        # write-checking on the same transcript (tsgpu_twist_write_check_prove)
        wclaims = np.zeros((2, 4), dtype=np.uint64)
        r3 = np.zeros((max(k + t, 1), 4, 4), dtype=np.uint64); f3 = np.zeros(4, dtype=np.uint64)
        r4 = np.zeros((max(t, 1), 4, 4), dtype=np.uint64); f4 = np.zeros(4, dtype=np.uint64)
        self.ctx.check(lib().tsgpu_twist_write_check_prove(self.ctx._h, _p(addr), _p(vals), _p(isw), C.c_size_t(addr.shape[0]), C.c_size_t(memory_size),
                                                           transcript._h, _p(wclaims), _p(r3), _p(f3), _p(r4), _p(f4)))
        return TwistMemoryCheck.Proof(claims, SumCheckProof(r1[:k + t], f1), SumCheckProof(r2[:t], f2), wclaims, SumCheckProof(r3[:k + t], f3),
                                      SumCheckProof(r4[:t], f4))

    def verify(self, trace: "MemoryTrace", proof: "TwistMemoryCheck.Proof", transcript: Transcript) -> bool:
        addr, vals, isw = trace.arrays()
        return self.verify_arrays(addr, vals, isw, trace.memory_size, proof, transcript)

    def verify_arrays(self, addresses, values, is_write, memory_size: int, proof: "TwistMemoryCheck.Proof", transcript: Transcript) -> bool:
        addr = np.ascontiguousarray(addresses, dtype=np.uint64).reshape(-1)
        vals = _fr(values) if len(values) else np.empty((0, 4), dtype=np.uint64)
        isw = np.ascontiguousarray(is_write, dtype=np.uint8).reshape(-1)
        claims = np.ascontiguousarray(proof.claims, dtype=np.uint64).reshape(2, 4)
        r1 = np.ascontiguousarray(proof.read_check.round_polynomials, dtype=np.uint64).reshape(-1, 4, 4)
        r2 = np.ascontiguousarray(proof.val_evaluation.round_polynomials, dtype=np.uint64).reshape(-1, 4, 4)
        f1 = _fr(proof.read_check.final_evaluation, 1); f2 = _fr(proof.val_evaluation.final_evaluation, 1)
        ok = C.c_int(0)
        self.ctx.check(lib().tsgpu_twist_memory_check_verify(self.ctx._h, _p(addr), _p(vals), _p(isw), C.c_size_t(addr.shape[0]), C.c_size_t(memory_size),
                                                             transcript._h, _p(claims), _p(r1), C.c_size_t(r1.shape[0]), _p(f1),
                                                             _p(r2), C.c_size_t(r2.shape[0]), _p(f2), C.byref(ok)))
        if not ok.value or proof.write_check is None:
            return bool(ok.value)
        wclaims = np.ascontiguousarray(proof.write_claims, dtype=np.uint64).reshape(2, 4)
        r3 = np.ascontiguousarray(proof.write_check.round_polynomials, dtype=np.uint64).reshape(-1, 4, 4)
        r4 = np.ascontiguousarray(proof.write_val_evaluation.round_polynomials, dtype=np.uint64).reshape(-1, 4, 4)
        f3 = _fr(proof.write_check.final_evaluation, 1); f4 = _fr(proof.write_val_evaluation.final_evaluation, 1)
        self.ctx.check(lib().tsgpu_twist_write_check_verify(self.ctx._h, _p(addr), _p(vals), _p(isw), C.c_size_t(addr.shape[0]), C.c_size_t(memory_size),
                                                            transcript._h, _p(wclaims), _p(r3), C.c_size_t(r3.shape[0]), _p(f3),
                                                            _p(r4), C.c_size_t(r4.shape[0]), _p(f4), C.byref(ok)))
        return bool(ok.value)

    # ---- binding to the KZG commitments of a Twist proof (tsgpu_transcript_bind_proof / tsgpu_twist_commitments_match)
    @staticmethod
    def bind(transcript: Transcript, proof: "TwistProof") -> Transcript:
        """absorb the proof's address / value commitment hashes (labels of src/twist.rs:157-160) before the sum-checks draw anything; returns the transcript"""
        rc = lib().tsgpu_transcript_bind_proof(transcript._h, proof._h, C.c_int(0))
        assert rc == 0, rc
        return transcript

    def commitments_match(self, params, proof: "TwistProof", trace: "MemoryTrace") -> bool:
        """do the proof's two commitments commit to THIS trace's addresses and values?  (recomputed on the device from the clear statement)"""
        addr, vals, _ = trace.arrays()
        addr = np.ascontiguousarray(addr, dtype=np.uint64).reshape(-1)
        vals = _fr(vals) if len(vals) else np.empty((0, 4), dtype=np.uint64)
        ok = C.c_int(0)
        self.ctx.check(lib().tsgpu_twist_commitments_match(self.ctx._h, params._h, proof._h, _p(addr), _p(vals), C.c_size_t(addr.shape[0]), C.byref(ok)))
        return bool(ok.value)


# ---------------------------------------------------------------------------------------------------------------------------------
# Small host-side helpers of the reference that are not on the GPU path (src/utils.rs:207-269 `field_utils`, src/polynomials.rs:296-371
# `poly_utils`, src/commitments.rs:317-375 `polynomial_division`): mirrored so that the reference's own tests can be restated one for one.
# Plain Python integers on canonical values; only lagrange_interpolate on the nodes 0..n-1 and inner_product of tables run on the device.
R_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_R2 = pow(1 << 256, 2, R_MODULUS)


def fe_from_int(x: int) -> np.ndarray:
    """canonical integer (any size, reduced mod r) -> FieldElement"""
    raw = np.array([[(x % R_MODULUS) >> (64 * i) & 0xFFFFFFFFFFFFFFFF for i in range(4)]], dtype=np.uint64)
    r2 = np.array([[_R2 >> (64 * i) & 0xFFFFFFFFFFFFFFFF for i in range(4)]], dtype=np.uint64)
    out = np.empty((1, 4), dtype=np.uint64)
    lib().tsgpu_fr_mul(_p(raw), _p(r2), _p(out))          # raw * R^2 / R = raw * R: the Montgomery form
    return out[0]


def fe_add(a, b) -> np.ndarray:
    a = _fr(a, 1); b = _fr(b, 1); out = np.empty((1, 4), dtype=np.uint64)
    lib().tsgpu_fr_add(_p(a), _p(b), _p(out))
    return out[0]


def fe_mul(a, b) -> np.ndarray:
    a = _fr(a, 1); b = _fr(b, 1); out = np.empty((1, 4), dtype=np.uint64)
    lib().tsgpu_fr_mul(_p(a), _p(b), _p(out))
    return out[0]


def fe_inverse(a) -> np.ndarray:
    return fe_from_int(pow(fe_to_int(a), -1, R_MODULUS))


class field_utils:                                       # noqa: N801 - the reference's module name
    """src/utils.rs:207-269"""

    @staticmethod
    def inner_product(a, b) -> np.ndarray:
        a = _fr(a); b = _fr(b)
        assert a.shape[0] == b.shape[0], "Vector lengths must match"
        return fe_from_int(sum(fe_to_int(x) * fe_to_int(y) for x, y in zip(a, b)))

    @staticmethod
    def horner_eval(coeffs, point) -> np.ndarray:
        coeffs = _fr(coeffs) if len(coeffs) else np.empty((0, 4), dtype=np.uint64)
        point = _fr(point, 1); out = np.empty((1, 4), dtype=np.uint64)
        lib().tsgpu_horner_eval(_p(coeffs), C.c_size_t(coeffs.shape[0]), _p(point), _p(out))
        return out[0]

    @staticmethod
    def powers(x, n: int) -> np.ndarray:
        xi, cur, out = fe_to_int(x), 1, []
        for _ in range(n):
            out.append(fe_from_int(cur)); cur = cur * xi % R_MODULUS
        return np.stack(out) if out else np.empty((0, 4), dtype=np.uint64)

    @staticmethod
    def vanishing_poly_eval(points, point) -> np.ndarray:
        z, acc = fe_to_int(point), 1
        for s in _fr(points):
            acc = acc * (z - fe_to_int(s)) % R_MODULUS
        return fe_from_int(acc)

    @staticmethod
    def batch_inverse(elements) -> np.ndarray:
        els = [fe_to_int(e) for e in _fr(elements)] if len(elements) else []
        return np.stack([fe_from_int(pow(e, -1, R_MODULUS)) for e in els]) if els else np.empty((0, 4), dtype=np.uint64)


class poly_utils:                                        # noqa: N801
    """src/polynomials.rs:296-371"""

    @staticmethod
    def lagrange_interpolate(ctx: Context, points) -> np.ndarray:
        """points = [(x_i, y_i)]; on the nodes x_i = i (the only use on the prove path, src/twist.rs:307-315) the device interpolation runs,
        any other node set takes the reference's own O(n^2) formula on the host"""
        xs = [fe_to_int(x) for x, _ in points]
        ys = np.stack([np.asarray(y, dtype=np.uint64).reshape(4) for _, y in points]) if points else np.empty((0, 4), dtype=np.uint64)
        if xs == list(range(len(xs))):
            return ctx.interpolate_iota(ys) if xs else ys
        n, yi, p = len(xs), [fe_to_int(y) for y in ys], R_MODULUS
        coeffs = [0] * n
        for i in range(n):
            basis, denom = [1], 1
            for j in range(n):
                if j != i:
                    basis = [(a - xs[j] * b) % p for a, b in zip([0] + basis, basis + [0])]
                    denom = denom * (xs[i] - xs[j]) % p
            scale = yi[i] * pow(denom, -1, p) % p
            coeffs = [(c + scale * b) % p for c, b in zip(coeffs, basis)]
        return np.stack([fe_from_int(c) for c in coeffs])

    @staticmethod
    def evaluate_polynomial(coeffs, point) -> np.ndarray:
        return field_utils.horner_eval(coeffs, point)

    @staticmethod
    def derivative(coeffs) -> np.ndarray:
        cs = [fe_to_int(c) for c in _fr(coeffs)] if len(coeffs) else []
        if len(cs) <= 1:
            return fe(0).reshape(1, 4)
        return np.stack([fe_from_int(c * i) for i, c in enumerate(cs) if i >= 1])


def polynomial_division(dividend, divisor) -> np.ndarray:
    """src/commitments.rs:317-375: quotient of the long division; TwistAndShoutError(Polynomial, "Cannot divide by zero polynomial")"""
    p = R_MODULUS
    a = [fe_to_int(c) for c in _fr(dividend)] if len(dividend) else []
    b = [fe_to_int(c) for c in _fr(divisor)] if len(divisor) else []
    while b and b[-1] == 0:
        b.pop()
    if not b:
        raise TwistAndShoutError(5, "Cannot divide by zero polynomial")
    if len(a) < len(b):
        return np.empty((0, 4), dtype=np.uint64)
    q = [0] * (len(a) - len(b) + 1)
    inv = pow(b[-1], -1, p)
    for k in range(len(q) - 1, -1, -1):
        q[k] = a[k + len(b) - 1] * inv % p
        for j, bj in enumerate(b):
            a[k + j] = (a[k + j] - q[k] * bj) % p
    return np.stack([fe_from_int(c) for c in q])
