"""ctypes binding of libtsgpu.so (include/tsgpu.h) - the only way Python code reaches the CUDA path.

There is deliberately no fallback: if the shared library is missing, or no CUDA device is present,
the calls raise.  Numpy arrays carry the reference layouts unchanged:
    Fr  : uint64[..., 4]   Montgomery limbs (ark_bn254::Fr)
    G1  : uint64[..., 12]  Jacobian {x, y, z}
    G1a : uint64[..., 8]   affine {x, y}
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TSGPU_LIB") or os.path.join(_HERE, "libtsgpu.so")   # TSGPU_LIB: tuning builds only

# TwistAndShoutError variants (reference src/lib.rs:59-78)
ERROR_NAMES = {1: "InvalidParameters", 2: "ProofGeneration", 3: "ProofVerification", 4: "Commitment",
               5: "Polynomial", 6: "SumCheck"}


class TwistAndShoutError(Exception):
    """Mirror of the reference error enum: `.variant` is the Rust variant name, str() the message."""

    def __init__(self, code: int, message: str):
        self.code = code
        self.variant = ERROR_NAMES.get(code, f"Unknown({code})")
        super().__init__(f"{self.variant}: {message}")
        self.message = message


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.tsgpu_last_error.restype = C.c_char_p
        L.tsgpu_last_error.argtypes = [C.c_void_p]
        L.tsgpu_launch_count.restype = C.c_uint64
        L.tsgpu_launch_count.argtypes = [C.c_void_p]
        L.tsgpu_counter_read.restype = C.c_uint64
        L.tsgpu_counter_read.argtypes = [C.c_void_p, C.c_char_p]
        L.tsgpu_table_num_vars.restype = C.c_uint
        L.tsgpu_table_num_vars.argtypes = [C.c_void_p]
        L.tsgpu_sc_num_vars.restype = C.c_uint
        L.tsgpu_sc_num_vars.argtypes = [C.c_void_p]
        L.tsgpu_destroy.argtypes = [C.c_void_p]
        L.tsgpu_table_free.argtypes = [C.c_void_p, C.c_void_p]
        L.tsgpu_sc_end.argtypes = [C.c_void_p]
        L.tsgpu_srs_len.restype = C.c_size_t
        L.tsgpu_srs_len.argtypes = [C.c_void_p]
        L.tsgpu_srs_free.argtypes = [C.c_void_p, C.c_void_p]
        L.tsgpu_srs_can_lagrange.argtypes = [C.c_void_p]
        L.tsgpu_srs_has_lagrange.argtypes = [C.c_void_p, C.c_size_t]
        L.tsgpu_poly_len.restype = C.c_size_t
        L.tsgpu_poly_len.argtypes = [C.c_void_p]
        L.tsgpu_poly_free.argtypes = [C.c_void_p, C.c_void_p]
        L.tsgpu_g1_hash.argtypes = [C.c_void_p, C.c_void_p]
        L.tsgpu_g1_compress.argtypes = [C.c_void_p, C.c_void_p]
        L.tsgpu_g1_equal.argtypes = [C.c_void_p, C.c_void_p]
        L.tsgpu_timer_reset.argtypes = [C.c_void_p]
        L.tsgpu_params_free.argtypes = [C.c_void_p, C.c_void_p]
        L.tsgpu_kzg_verify.argtypes = [C.c_void_p] * 5 + [C.c_void_p]
        L.tsgpu_kzg_batch_verify.argtypes = [C.c_void_p] * 5 + [C.c_size_t, C.c_void_p]
        L.tsgpu_transcript_new.restype = C.c_void_p
        L.tsgpu_transcript_new.argtypes = [C.c_void_p]
        L.tsgpu_transcript_free.argtypes = [C.c_void_p]
        L.tsgpu_transcript_append.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.tsgpu_transcript_challenge.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p]
        L.tsgpu_transcript_state_len.restype = C.c_size_t
        L.tsgpu_transcript_state_len.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _p(a: np.ndarray) -> C.c_void_p:
    return a.ctypes.data_as(C.c_void_p)


def _fr(a, n: Optional[int] = None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    if n is not None and a.shape[0] != n:
        raise ValueError(f"expected {n} field elements, got {a.shape[0]}")
    return a


class Context:
    """tsgpu_ctx: one GPU + one stream.  `stream` may be a raw cudaStream_t (int) such as
    torch.cuda.current_stream().cuda_stream so that torch events time the library's kernels."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._h = C.c_void_p()
        rc = lib().tsgpu_init(C.c_int(device), C.c_void_p(stream or 0), C.byref(self._h))
        if rc:
            raise TwistAndShoutError(rc, "tsgpu_init failed: no usable CUDA device (there is no CPU fallback)")
        self.device = device

    def close(self):
        if self._h:
            lib().tsgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc:
            raise TwistAndShoutError(rc, lib().tsgpu_last_error(self._h).decode())

    @property
    def launch_count(self) -> int:
        return int(lib().tsgpu_launch_count(self._h))

    @property
    def sm_count(self) -> int:
        return int(lib().tsgpu_sm_count(self._h))

    # ---- multi-GPU: one process per GPU, NCCL communicator inside the library (csrc/comm.cu)
    def comm_init(self, nranks: int, rank: int, unique_id: Optional[bytes] = None):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        self.check(lib().tsgpu_comm_init(self._h, C.c_int(nranks), C.c_int(rank), buf))

    def comm_init_torch(self, group=None):
        """create the library communicator over the ranks of an initialised torch.distributed group: rank 0 draws the
        NCCL unique id, torch.distributed (any backend) carries its 128 bytes to the other ranks"""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        ids = [unique_id() if rank == 0 and world > 1 else None]
        if world > 1:
            dist.broadcast_object_list(ids, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self.comm_init(world, rank, ids[0])

    @property
    def comm_size(self) -> int:
        return int(lib().tsgpu_comm_size(self._h))

    @property
    def comm_peer_exchange(self) -> bool:
        """True when the ranks exchange through peer-mapped mailboxes (round sums inside the round kernel) instead of NCCL collectives"""
        return bool(lib().tsgpu_comm_peer_exchange(self._h))

    @property
    def comm_rank(self) -> int:
        return int(lib().tsgpu_comm_rank(self._h))

    def comm_allgather(self, x: np.ndarray) -> np.ndarray:
        """(m,) uint64 per rank -> (ranks, m) on every rank"""
        x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1)
        out = np.empty((self.comm_size, x.shape[0]), dtype=np.uint64)
        self.check(lib().tsgpu_comm_allgather(self._h, _p(x), C.c_size_t(x.nbytes), _p(out)))
        return out

    def counter(self, name: str) -> int:
        """work counters: launches, msm_calls, msm_points, msm_entries"""
        return int(lib().tsgpu_counter_read(self._h, name.encode()))

    def set_tuning(self, key: str, value: int):
        self.check(lib().tsgpu_set_tuning(self._h, key.encode(), C.c_long(value)))

    def timer_read(self, name: str):
        ms = C.c_double(0); cnt = C.c_uint64(0)
        self.check(lib().tsgpu_timer_read(self._h, name.encode(), C.byref(ms), C.byref(cnt)))
        return ms.value, int(cnt.value)

    def timer_reset(self):
        lib().tsgpu_timer_reset(self._h)

    def synchronize(self):
        self.check(lib().tsgpu_synchronize(self._h))

    # ---- tables
    def table_upload(self, evals, num_vars: Optional[int] = None) -> "Table":
        evals = _fr(evals)
        n = evals.shape[0]
        if num_vars is None:
            num_vars = max(n.bit_length() - 1, 0)
            if n == 0 or (1 << num_vars) != n:
                raise ValueError("Evaluation vector length must be a power of 2")   # polynomials.rs:30
        h = C.c_void_p()
        self.check(lib().tsgpu_table_upload(self._h, _p(evals), C.c_size_t(n), C.c_uint(num_vars), C.byref(h)))
        return Table(self, h)

    def table_eq(self, w) -> "Table":
        w = _fr(w)
        h = C.c_void_p()
        self.check(lib().tsgpu_table_eq(self._h, _p(w), C.c_uint(w.shape[0]), C.byref(h)))
        return Table(self, h)

    def table_one_hot_rows(self, idx, log_k: int, num_vars: int) -> "Table":
        idx = np.ascontiguousarray(idx, dtype=np.uint64).reshape(-1)
        h = C.c_void_p()
        self.check(lib().tsgpu_table_one_hot_rows(self._h, _p(idx), C.c_size_t(idx.shape[0]), C.c_uint(log_k),
                                                  C.c_uint(num_vars), C.byref(h)))
        return Table(self, h)

    def table_from_u64(self, v, num_vars: int) -> "Table":
        v = np.ascontiguousarray(v, dtype=np.uint64).reshape(-1)
        h = C.c_void_p()
        self.check(lib().tsgpu_table_from_u64(self._h, _p(v), C.c_size_t(v.shape[0]), C.c_uint(num_vars), C.byref(h)))
        return Table(self, h)

    def table_one_hot(self, num_vars: int, index: int) -> "Table":
        """MultilinearExtension::one_hot (polynomials.rs:71-82)"""
        h = C.c_void_p()
        self.check(lib().tsgpu_table_one_hot(self._h, C.c_uint(num_vars), C.c_size_t(index), C.byref(h)))
        return Table(self, h)

    def table_from_sparse(self, num_vars: int, sparse_entries) -> "Table":
        """MultilinearExtension::from_sparse (polynomials.rs:52-67): sparse_entries = [(index, value)], a repeated index keeps its last value"""
        idx = np.ascontiguousarray([int(i) for i, _ in sparse_entries], dtype=np.uint64).reshape(-1)
        vals = _fr(np.stack([np.asarray(v, dtype=np.uint64).reshape(4) for _, v in sparse_entries])) if len(sparse_entries) else np.empty((0, 4), dtype=np.uint64)
        h = C.c_void_p()
        self.check(lib().tsgpu_table_from_sparse(self._h, C.c_uint(num_vars), _p(idx), _p(vals), C.c_size_t(idx.shape[0]), C.byref(h)))
        return Table(self, h)

    def table_less_than(self, num_vars: int) -> "Table":
        """LessThanPolynomial::new(num_vars).to_multilinear_extension() (polynomials.rs:243-263), generated on the device"""
        h = C.c_void_p()
        self.check(lib().tsgpu_table_less_than(self._h, C.c_uint(num_vars), C.byref(h)))
        return Table(self, h)

    # ---- host-buffer MLE calls
    def mle_evaluate(self, evals, point) -> np.ndarray:
        evals = _fr(evals); nv = evals.shape[0].bit_length() - 1
        point = _fr(point, nv)
        out = np.empty(4, dtype=np.uint64)
        self.check(lib().tsgpu_mle_evaluate(self._h, _p(evals), C.c_uint(nv), _p(point), _p(out)))
        return out

    def mle_partial_evaluate(self, evals, fixed) -> np.ndarray:
        evals = _fr(evals); nv = evals.shape[0].bit_length() - 1
        fixed = _fr(fixed); k = fixed.shape[0]
        out = np.empty((1 << max(nv - k, 0), 4), dtype=np.uint64)
        self.check(lib().tsgpu_mle_partial_evaluate(self._h, _p(evals), C.c_uint(nv), _p(fixed), C.c_uint(k), _p(out)))
        return out

    def sumcheck(self, tables: Sequence["Table"]) -> "SumCheckRounds":
        return SumCheckRounds(self, tables)

    # ---- SRS / KZG
    def srs_generate(self, tau, n: int) -> "Srs":
        h = C.c_void_p()
        self.check(lib().tsgpu_srs_generate(self._h, _p(_fr(tau, 1)), C.c_size_t(n), C.byref(h)))
        return Srs(self, h)

    def srs_generate_range(self, tau, first: int, n: int) -> "Srs":
        """g1_powers[first .. first + n) only (the slice a point-sharded MSM rank holds)"""
        h = C.c_void_p()
        self.check(lib().tsgpu_srs_generate_range(self._h, _p(_fr(tau, 1)), C.c_size_t(first), C.c_size_t(n), C.byref(h)))
        return Srs(self, h)

    def srs_upload(self, powers_jac) -> "Srs":
        powers_jac = np.ascontiguousarray(powers_jac, dtype=np.uint64).reshape(-1, 12)
        h = C.c_void_p()
        self.check(lib().tsgpu_srs_upload(self._h, _p(powers_jac), C.c_size_t(powers_jac.shape[0]), C.byref(h)))
        return Srs(self, h)

    def poly_upload(self, coeffs) -> "Poly":
        coeffs = _fr(coeffs)
        h = C.c_void_p()
        self.check(lib().tsgpu_poly_upload(self._h, _p(coeffs), C.c_size_t(coeffs.shape[0]), C.byref(h)))
        return Poly(self, h)

    def poly_from_u64(self, v, padded: Optional[int] = None) -> "Poly":
        v = np.ascontiguousarray(v, dtype=np.uint64).reshape(-1)
        h = C.c_void_p()
        self.check(lib().tsgpu_poly_from_u64(self._h, _p(v), C.c_size_t(v.shape[0]), C.c_size_t(padded or v.shape[0]), C.byref(h)))
        return Poly(self, h)

    def poly_upload_padded(self, vals, padded: int) -> "Poly":
        vals = _fr(vals)
        h = C.c_void_p()
        self.check(lib().tsgpu_poly_upload_padded(self._h, _p(vals), C.c_size_t(vals.shape[0]), C.c_size_t(padded), C.byref(h)))
        return Poly(self, h)

    def interpolate_prepare(self, log_n: int):
        self.check(lib().tsgpu_interpolate_prepare(self._h, C.c_uint(log_n)))

    def interpolate_iota(self, values) -> np.ndarray:
        """poly_utils::lagrange_interpolate on the points (i, values[i]) (src/polynomials.rs:301-352)"""
        values = _fr(values)
        out = np.empty_like(values)
        self.check(lib().tsgpu_interpolate_iota(self._h, _p(values), C.c_size_t(values.shape[0]), _p(out)))
        return out

    def msm_g1(self, bases_affine, scalars) -> np.ndarray:
        bases_affine = np.ascontiguousarray(bases_affine, dtype=np.uint64).reshape(-1, 8)
        scalars = _fr(scalars)
        if bases_affine.shape[0] < scalars.shape[0]:
            raise ValueError("fewer bases than scalars")
        out = np.empty(12, dtype=np.uint64)
        self.check(lib().tsgpu_msm_g1(self._h, _p(bases_affine), _p(scalars), C.c_size_t(scalars.shape[0]), _p(out)))
        return out


class Table:
    """tsgpu_table: MultilinearExtension.evaluations resident in HBM."""

    def __init__(self, ctx: Context, handle: C.c_void_p):
        self.ctx = ctx
        self._h = handle

    @property
    def num_vars(self) -> int:
        return int(lib().tsgpu_table_num_vars(self._h))

    def download(self) -> np.ndarray:
        out = np.empty((1 << self.num_vars, 4), dtype=np.uint64)
        self.ctx.check(lib().tsgpu_table_download(self.ctx._h, self._h, _p(out)))
        return out

    def clone(self) -> "Table":
        h = C.c_void_p()
        self.ctx.check(lib().tsgpu_table_clone(self.ctx._h, self._h, C.byref(h)))
        return Table(self.ctx, h)

    def evaluate(self, point) -> np.ndarray:
        point = _fr(point, self.num_vars)
        out = np.empty(4, dtype=np.uint64)
        self.ctx.check(lib().tsgpu_table_evaluate(self.ctx._h, self._h, _p(point), _p(out)))
        return out

    def evaluate_sharded(self, num_vars: int, point) -> np.ndarray:
        """this table is the local slice (high index bits = rank) of a num_vars-variable MLE spread over ctx's communicator"""
        point = _fr(point, num_vars)
        out = np.empty(4, dtype=np.uint64)
        self.ctx.check(lib().tsgpu_table_evaluate_sharded(self.ctx._h, self._h, C.c_uint(num_vars), _p(point), _p(out)))
        return out

    def partial_evaluate(self, fixed) -> "Table":
        fixed = _fr(fixed)
        h = C.c_void_p()
        self.ctx.check(lib().tsgpu_table_partial_evaluate(self.ctx._h, self._h, _p(fixed), C.c_uint(fixed.shape[0]), C.byref(h)))
        return Table(self.ctx, h)

    def bind(self, r):
        r = _fr(r, 1)
        self.ctx.check(lib().tsgpu_table_bind(self.ctx._h, self._h, _p(r)))

    def add(self, other: "Table") -> "Table":
        """MultilinearExtension::add (polynomials.rs:164-176)"""
        h = C.c_void_p()
        self.ctx.check(lib().tsgpu_table_add(self.ctx._h, self._h, other._h, C.byref(h)))
        return Table(self.ctx, h)

    def scalar_mul(self, scalar) -> "Table":
        """MultilinearExtension::scalar_mul (polynomials.rs:179-189)"""
        scalar = _fr(scalar, 1)
        h = C.c_void_p()
        self.ctx.check(lib().tsgpu_table_scalar_mul(self.ctx._h, self._h, _p(scalar), C.byref(h)))
        return Table(self.ctx, h)

    def scatter_add(self, idx, log_k: int) -> "Table":
        """out[x] = sum over j with idx[j] == x of self[j]: the one-hot lookup matrix applied to this weight vector"""
        idx = np.ascontiguousarray(idx, dtype=np.uint64).reshape(-1)
        h = C.c_void_p()
        self.ctx.check(lib().tsgpu_table_scatter_add(self.ctx._h, self._h, _p(idx), C.c_size_t(idx.shape[0]), C.c_uint(log_k), C.byref(h)))
        return Table(self.ctx, h)

    def gather(self, idx, num_vars: int) -> "Table":
        """out[j] = self[idx[j]], zero padded to 2^num_vars entries"""
        idx = np.ascontiguousarray(idx, dtype=np.uint64).reshape(-1)
        h = C.c_void_p()
        self.ctx.check(lib().tsgpu_table_gather(self.ctx._h, self._h, _p(idx), C.c_size_t(idx.shape[0]), C.c_uint(num_vars), C.byref(h)))
        return Table(self.ctx, h)

    def inner_product(self, other: "Table") -> np.ndarray:
        """field_utils::inner_product (utils.rs:210-213)"""
        out = np.empty(4, dtype=np.uint64)
        self.ctx.check(lib().tsgpu_table_inner_product(self.ctx._h, self._h, other._h, _p(out)))
        return out

    def sum_evaluations(self) -> np.ndarray:
        """MultilinearExtension::sum_evaluations (polynomials.rs:192-195)"""
        out = np.empty(4, dtype=np.uint64)
        self.ctx.check(lib().tsgpu_table_sum_evaluations(self.ctx._h, self._h, _p(out)))
        return out


    def free(self):
        if self._h:
            lib().tsgpu_table_free(self.ctx._h, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass


class SumCheckRounds:
    """tsgpu_sc: round-stepped sum-check prover over a product of 1..3 tables (consumed)."""

    def __init__(self, ctx: Context, tables: Sequence[Table]):
        self.ctx = ctx
        self.tables = list(tables)
        arr = (C.c_void_p * len(self.tables))(*[t._h for t in self.tables])
        self._h = C.c_void_p()
        ctx.check(lib().tsgpu_sc_begin(ctx._h, arr, C.c_int(len(self.tables)), C.byref(self._h)))

    @property
    def vars_left(self) -> int:
        return int(lib().tsgpu_sc_num_vars(self._h))

    def round_eval(self) -> np.ndarray:
        out = np.empty((4, 4), dtype=np.uint64)
        self.ctx.check(lib().tsgpu_sc_round_eval(self._h, _p(out)))
        return out

    def bind(self, r):
        r = _fr(r, 1)
        self.ctx.check(lib().tsgpu_sc_bind(self._h, _p(r)))

    def exclusive(self, on: bool = True):
        """promise (on) that nothing else is enqueued on the context until these rounds end: lets the small d = 2 claim-form rounds run in the persistent tail kernel"""
        self.ctx.check(lib().tsgpu_sc_exclusive(self._h, C.c_int(1 if on else 0)))

    def bind_eval(self, r, claim=None) -> np.ndarray:
        """fused bind(r) + evaluation of the next round; with `claim` (= g(r) of the round just bound) g(1) is derived as claim - g(0)"""
        r = _fr(r, 1)
        out = np.empty((4, 4), dtype=np.uint64)
        if claim is None:
            self.ctx.check(lib().tsgpu_sc_bind_eval(self._h, _p(r), _p(out)))
        else:
            claim = _fr(claim, 1)
            self.ctx.check(lib().tsgpu_sc_bind_eval_claim(self._h, _p(r), _p(claim), _p(out)))
        return out

    def final(self) -> np.ndarray:
        out = np.empty((len(self.tables), 4), dtype=np.uint64)
        self.ctx.check(lib().tsgpu_sc_final(self._h, _p(out)))
        return out

    def end(self):
        if self._h:
            lib().tsgpu_sc_end(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.end()
        except Exception:
            pass


class Transcript:
    """Host-side Fiat-Shamir transcript of the reference (src/utils.rs:134-204); no GPU needed."""

    def __init__(self, seed: bytes = b"\0" * 32):
        buf = C.create_string_buffer(bytes(seed), 32)
        self._h = C.c_void_p(lib().tsgpu_transcript_new(buf))

    def append_field_element(self, label: bytes, x):
        self.append_field_elements(label, _fr(x, 1))

    def append_field_elements(self, label: bytes, xs):
        xs = _fr(xs)
        lib().tsgpu_transcript_append(self._h, label, len(label), _p(xs), xs.shape[0])

    def challenge_field_element(self, label: bytes) -> np.ndarray:
        out = np.empty(4, dtype=np.uint64)
        lib().tsgpu_transcript_challenge(self._h, label, len(label), _p(out))
        return out

    def challenge_field_elements(self, label: bytes, count: int) -> np.ndarray:
        """labels "{label}_{i}" (utils.rs:195-203)"""
        if count == 0:
            return np.empty((0, 4), dtype=np.uint64)
        return np.stack([self.challenge_field_element(label + b"_" + str(i).encode()) for i in range(count)])

    @property
    def state_len(self) -> int:
        return int(lib().tsgpu_transcript_state_len(self._h))

    def __del__(self):
        try:
            if self._h:
                lib().tsgpu_transcript_free(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


class SumCheckProof:
    """src/sumcheck.rs:25-31"""

    def __init__(self, round_polynomials: np.ndarray, final_evaluation: np.ndarray):
        self.round_polynomials = round_polynomials     # (num_vars, 4, 4) coefficients low -> high
        self.final_evaluation = final_evaluation       # (4,)


class SumCheck:
    """Mirror of the reference's SumCheck { num_vars, claimed_sum } (src/sumcheck.rs:15-53) with the structured
    prover: `prove_product(tables, transcript)` equals `prove(|v| prod_t mle_t.evaluate(v), transcript)`."""

    def __init__(self, num_vars: int, claimed_sum):
        self.num_vars = num_vars
        self.claimed_sum = _fr(claimed_sum, 1).reshape(4)

    def prove_product(self, ctx: Context, tables: Sequence[Table], transcript: Transcript, return_aux: bool = False):
        tables = list(tables)
        if any(t.num_vars != self.num_vars for t in tables):
            raise TwistAndShoutError(1, "Number of variables must match")
        nv, d = self.num_vars, len(tables)
        arr = (C.c_void_p * d)(*[t._h for t in tables])
        rp = np.zeros((max(nv, 1), 4, 4), dtype=np.uint64)
        fe = np.zeros(4, dtype=np.uint64)
        ch = np.zeros((max(nv, 1), 4), dtype=np.uint64)
        fin = np.zeros((d, 4), dtype=np.uint64)
        ctx.check(lib().tsgpu_sumcheck_prove_product(ctx._h, arr, C.c_int(d), _p(self.claimed_sum), transcript._h,
                                                     _p(rp), _p(fe), _p(ch), _p(fin)))
        proof = SumCheckProof(rp[:nv], fe)
        return (proof, ch[:nv], fin) if return_aux else proof

    def prove_product_sharded(self, ctx: Context, local_tables: Sequence[Table], transcript: Transcript, return_aux: bool = False):
        """same proof with the hypercube sliced over the ranks of ctx's communicator (ctx.comm_init*): `local_tables` are this
        rank's slices (high index bits = rank), num_vars - log2(ranks) variables each; one 256-byte all-reduce per round"""
        tables = list(local_tables)
        nv, d = self.num_vars, len(tables)
        arr = (C.c_void_p * d)(*[t._h for t in tables])
        rp = np.zeros((max(nv, 1), 4, 4), dtype=np.uint64)
        fe = np.zeros(4, dtype=np.uint64)
        ch = np.zeros((max(nv, 1), 4), dtype=np.uint64)
        fin = np.zeros((d, 4), dtype=np.uint64)
        ctx.check(lib().tsgpu_sumcheck_prove_product_sharded(ctx._h, arr, C.c_int(d), C.c_uint(nv), _p(self.claimed_sum), transcript._h,
                                                             _p(rp), _p(fe), _p(ch), _p(fin)))
        proof = SumCheckProof(rp[:nv], fe)
        return (proof, ch[:nv], fin) if return_aux else proof

    def verify(self, proof: SumCheckProof, transcript: Transcript):
        """-> (is_valid, challenges); raises SumCheck("Proof has wrong number of rounds") like sumcheck.rs:118-122"""
        rp = np.ascontiguousarray(proof.round_polynomials, dtype=np.uint64).reshape(-1, 4, 4)
        ch = np.zeros((max(rp.shape[0], 1), 4), dtype=np.uint64)
        valid = C.c_int(0)
        rc = lib().tsgpu_sumcheck_verify(C.c_uint(self.num_vars), _p(self.claimed_sum), _p(rp), C.c_size_t(rp.shape[0]),
                                         _p(_fr(proof.final_evaluation, 1)), transcript._h, C.byref(valid), _p(ch))
        if rc:
            raise TwistAndShoutError(rc, "Proof has wrong number of rounds")
        return bool(valid.value), ch[:rp.shape[0]]


class Srs:
    """tsgpu_srs: CommitmentParams.g1_powers resident in HBM."""

    def __init__(self, ctx: Context, handle: C.c_void_p):
        self.ctx = ctx
        self._h = handle

    def __len__(self) -> int:
        return int(lib().tsgpu_srs_len(self._h))

    def download(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = len(self) - first if count is None else count
        out = np.empty((count, 12), dtype=np.uint64)
        self.ctx.check(lib().tsgpu_srs_download(self.ctx._h, self._h, C.c_size_t(first), C.c_size_t(count), _p(out)))
        return out

    def can_lagrange(self) -> bool:
        """True when the handle keeps the trapdoor (made by srs_generate / setup_params): evaluation-basis commits possible"""
        return bool(lib().tsgpu_srs_can_lagrange(self._h))

    def has_lagrange(self, m: int) -> bool:
        return bool(lib().tsgpu_srs_has_lagrange(self._h, C.c_size_t(m)))

    def lagrange_prepare(self, m: int):
        """build [L_j(tau)]_1 for the nodes 0..m-1 (cached in the handle)"""
        self.ctx.check(lib().tsgpu_srs_lagrange_prepare(self.ctx._h, self._h, C.c_size_t(m)))

    def free(self):
        if self._h:
            lib().tsgpu_srs_free(self.ctx._h, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass


class Poly:
    """tsgpu_poly: coefficient vector resident in HBM."""

    def __init__(self, ctx: Context, handle: C.c_void_p):
        self.ctx = ctx
        self._h = handle

    def __len__(self) -> int:
        return int(lib().tsgpu_poly_len(self._h))

    def download(self) -> np.ndarray:
        out = np.empty((len(self), 4), dtype=np.uint64)
        self.ctx.check(lib().tsgpu_poly_download(self.ctx._h, self._h, _p(out)))
        return out

    def clone(self) -> "Poly":
        h = C.c_void_p()
        self.ctx.check(lib().tsgpu_poly_clone(self.ctx._h, self._h, C.byref(h)))
        return Poly(self.ctx, h)

    def interpolate_iota(self):
        """in place: values at 0..n-1 -> monomial coefficients"""
        self.ctx.check(lib().tsgpu_poly_interpolate_iota(self.ctx._h, self._h))
        return self

    def free(self):
        if self._h:
            lib().tsgpu_poly_free(self.ctx._h, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass


class KZGCommitment:
    """Mirror of `impl CommitmentScheme for KZGCommitment` (src/commitments.rs:156-199): associated functions,
    no self.  `params` is an Srs handle (CommitmentParams.g1_powers on the device); polynomials are either
    host arrays (uint64[n,4]) or Poly handles."""

    @staticmethod
    def commit(params: Srs, polynomial) -> np.ndarray:
        ctx = params.ctx
        out = np.empty(12, dtype=np.uint64)
        if isinstance(polynomial, Poly):
            ctx.check(lib().tsgpu_kzg_commit_dev(ctx._h, params._h, polynomial._h, _p(out)))
        else:
            polynomial = _fr(polynomial)
            ctx.check(lib().tsgpu_kzg_commit(ctx._h, params._h, _p(polynomial), C.c_size_t(polynomial.shape[0]), _p(out)))
        return out

    @staticmethod
    def open(params: Srs, polynomial, point):
        """-> (value, proof)"""
        ctx = params.ctx
        value = np.empty(4, dtype=np.uint64); proof = np.empty(12, dtype=np.uint64)
        point = _fr(point, 1)
        if isinstance(polynomial, Poly):
            ctx.check(lib().tsgpu_kzg_open_dev(ctx._h, params._h, polynomial._h, _p(point), _p(value), _p(proof)))
        else:
            polynomial = _fr(polynomial)
            ctx.check(lib().tsgpu_kzg_open(ctx._h, params._h, _p(polynomial), C.c_size_t(polynomial.shape[0]), _p(point),
                                           _p(value), _p(proof)))
        return value, proof


    @staticmethod
    def commit_values(params: Srs, values: "Poly") -> np.ndarray:
        """commit(vector_to_polynomial(values)) as one MSM over the values (evaluation-basis SRS, csrc/lagrange.cu)"""
        ctx = params.ctx
        out = np.empty(12, dtype=np.uint64)
        ctx.check(lib().tsgpu_kzg_commit_values_dev(ctx._h, params._h, values._h, _p(out)))
        return out

    @staticmethod
    def open_values(params: Srs, values: "Poly", point):
        """open(vector_to_polynomial(values), point) without the coefficients -> (value, proof)"""
        ctx = params.ctx
        value = np.empty(4, dtype=np.uint64); proof = np.empty(12, dtype=np.uint64)
        point = _fr(point, 1)
        ctx.check(lib().tsgpu_kzg_open_values_dev(ctx._h, params._h, values._h, _p(point), _p(value), _p(proof)))
        return value, proof


def unique_id() -> bytes:
    """ncclGetUniqueId (to be created on rank 0 and handed to every rank's Context.comm_init)"""
    buf = (C.c_uint8 * 128)()
    if lib().tsgpu_comm_unique_id(buf):
        raise RuntimeError("NCCL is not available: cannot create a communicator id")
    return bytes(buf)


def chacha20_u64(seed: bytes, n: int) -> np.ndarray:
    """n outputs of ChaCha20Rng::from_seed(seed).next_u64() (the generator behind tau, the challenges and the seeded benchmark traces); CPU."""
    assert len(seed) == 32
    out = np.empty(n, dtype=np.uint64)
    lib().tsgpu_chacha20_u64(seed, C.c_size_t(n), _p(out))
    return out


def chacha20_fr_then_u64(seed: bytes, num_fr: int, num_u64: int = 0):
    """num_fr draws of Fr::rand then num_u64 draws of next_u64 from one ChaCha20Rng::from_seed(seed); CPU."""
    assert len(seed) == 32
    f = np.empty((num_fr, 4), dtype=np.uint64); u = np.empty(num_u64, dtype=np.uint64)
    lib().tsgpu_chacha20_fr_then_u64(seed, C.c_size_t(num_fr), _p(f), C.c_size_t(num_u64), _p(u))
    return f, u


def statement_digest(domain: bytes, header, segments) -> bytes:
    """the 32-byte binding digest the non-parity constraint sum-checks absorb first (host/statement_digest.hpp); CPU."""
    hdr = np.ascontiguousarray(header, dtype=np.uint64)
    segs = [np.frombuffer(bytes(s), dtype=np.uint8) if not isinstance(s, np.ndarray) else np.ascontiguousarray(s).view(np.uint8).reshape(-1) for s in segments]
    ptrs = (C.c_void_p * max(len(segs), 1))(*[s.ctypes.data if s.size else None for s in segs])
    lens = (C.c_size_t * max(len(segs), 1))(*[s.size for s in segs])
    out = np.empty(32, dtype=np.uint8)
    lib().tsgpu_statement_digest(domain, _p(hdr), C.c_size_t(hdr.size), ptrs, lens, C.c_size_t(len(segs)), _p(out))
    return out.tobytes()


def g1_hash(point) -> np.ndarray:
    """KZGCommitmentValue::hash (src/commitments.rs:73-84); CPU."""
    point = np.ascontiguousarray(point, dtype=np.uint64).reshape(12)
    out = np.empty(4, dtype=np.uint64)
    lib().tsgpu_g1_hash(_p(point), _p(out))
    return out


def g1_compress(point) -> bytes:
    point = np.ascontiguousarray(point, dtype=np.uint64).reshape(12)
    out = np.empty(32, dtype=np.uint8)
    lib().tsgpu_g1_compress(_p(point), _p(out))
    return out.tobytes()


def g1_equal(a, b) -> bool:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(12); b = np.ascontiguousarray(b, dtype=np.uint64).reshape(12)
    return bool(lib().tsgpu_g1_equal(_p(a), _p(b)))
