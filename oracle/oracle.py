"""ctypes front-end of the C++ CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs; never by the product package.

Conventions (identical to the product's C ABI so buffers can be compared byte for byte):
  Fr / Fq element : numpy uint64[4], Montgomery form, little-endian limbs
  G1 Jacobian     : numpy uint64[12]  {X, Y, Z}
  G1 affine       : numpy uint64[8]   {x, y}, identity = all zero
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
P_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "ff.hpp", "g1.hpp", "rng.hpp")]
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_siphash13.restype = C.c_uint64
        _lib.orc_tr_new.restype = C.c_void_p
        _lib.orc_tr_state_len.restype = C.c_size_t
        _lib.orc_proof_max_bytes.restype = C.c_size_t
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def ncpu() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ----------------------------------------------------------------- int <-> limbs
def int_to_limbs(x: int) -> np.ndarray:
    return np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def ints_to_limbs(xs: Sequence[int]) -> np.ndarray:
    out = np.zeros((len(xs), 4), dtype=np.uint64)
    for i, x in enumerate(xs):
        for j in range(4):
            out[i, j] = (x >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def limbs_to_ints(a: np.ndarray) -> List[int]:
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    return [int(r[0]) | (int(r[1]) << 64) | (int(r[2]) << 128) | (int(r[3]) << 192) for r in a]


def fr_from_ints(xs: Sequence[int]) -> np.ndarray:
    """canonical ints -> Montgomery limbs (n,4)"""
    can = ints_to_limbs([x % R_MOD for x in xs])
    out = np.empty_like(can)
    lib().orc_fr_from_canonical(_p(can), C.c_size_t(len(xs)), _p(out))
    return out


def fr_to_ints(a: np.ndarray) -> List[int]:
    a = _u64(a).reshape(-1, 4)
    out = np.empty_like(a)
    lib().orc_fr_to_canonical(_p(a), C.c_size_t(a.shape[0]), _p(out))
    return limbs_to_ints(out)


def fq_from_ints(xs: Sequence[int]) -> np.ndarray:
    can = ints_to_limbs([x % P_MOD for x in xs])
    out = np.empty_like(can)
    lib().orc_fq_from_canonical(_p(can), C.c_size_t(len(xs)), _p(out))
    return out


def fq_to_ints(a: np.ndarray) -> List[int]:
    a = _u64(a).reshape(-1, 4)
    out = np.empty_like(a)
    lib().orc_fq_to_canonical(_p(a), C.c_size_t(a.shape[0]), _p(out))
    return limbs_to_ints(out)


def fr_from_u64(v: np.ndarray) -> np.ndarray:
    v = _u64(v).reshape(-1)
    out = np.empty((v.shape[0], 4), dtype=np.uint64)
    lib().orc_fr_from_u64(_p(v), C.c_size_t(v.shape[0]), _p(out))
    return out


def field_binop(field: str, op: str, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = _u64(a).reshape(-1, 4); b = _u64(b).reshape(-1, 4)
    out = np.empty_like(a)
    lib().orc_field_binop(0 if field == "fr" else 1, {"add": 0, "sub": 1, "mul": 2}[op], _p(a), _p(b),
                          C.c_size_t(a.shape[0]), _p(out))
    return out


def fr_inverse(a: np.ndarray) -> np.ndarray:
    a = _u64(a).reshape(-1, 4)
    out = np.empty_like(a)
    lib().orc_fr_inverse(_p(a), C.c_size_t(a.shape[0]), _p(out))
    return out


# ----------------------------------------------------------------- rng / hash / transcript
def chacha_fr_rand(seed: bytes, n: int) -> np.ndarray:
    out = np.empty((n, 4), dtype=np.uint64)
    lib().orc_chacha_fr_rand(seed, C.c_size_t(n), _p(out))
    return out


def chacha_u64(seed: bytes, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint64)
    lib().orc_chacha_u64(seed, C.c_size_t(n), _p(out))
    return out


def chacha_fr_then_u64(seed: bytes, nfr: int, nu: int) -> Tuple[np.ndarray, np.ndarray]:
    f = np.empty((nfr, 4), dtype=np.uint64); u = np.empty(nu, dtype=np.uint64)
    lib().orc_chacha_fr_then_u64(seed, C.c_size_t(nfr), _p(f), C.c_size_t(nu), _p(u))
    return f, u


def siphash13(data: bytes) -> int:
    return int(lib().orc_siphash13(data, C.c_size_t(len(data))))


class Transcript:
    """src/utils.rs:134-204"""

    def __init__(self, seed: bytes = b"\0" * 32):
        self._h = C.c_void_p(lib().orc_tr_new())

    def __del__(self):
        try:
            lib().orc_tr_free(self._h)
        except Exception:
            pass

    def append_field_element(self, label: bytes, x: np.ndarray):
        self.append_field_elements(label, _u64(x).reshape(1, 4))

    def append_field_elements(self, label: bytes, xs: np.ndarray):
        xs = _u64(xs).reshape(-1, 4)
        lib().orc_tr_append(self._h, label, C.c_size_t(len(label)), _p(xs), C.c_size_t(xs.shape[0]))

    def challenge_field_element(self, label: bytes) -> np.ndarray:
        out = np.empty(4, dtype=np.uint64)
        lib().orc_tr_challenge(self._h, label, C.c_size_t(len(label)), _p(out))
        return out

    def challenge_field_elements(self, label: bytes, count: int) -> np.ndarray:
        return np.stack([self.challenge_field_element(label + b"_" + str(i).encode()) for i in range(count)]) \
            if count else np.empty((0, 4), dtype=np.uint64)


# ----------------------------------------------------------------- polynomials / MLE
def lagrange_interpolate(xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    xs = _u64(xs).reshape(-1, 4); ys = _u64(ys).reshape(-1, 4)
    out = np.empty_like(ys)
    lib().orc_lagrange_interpolate(_p(xs), _p(ys), C.c_size_t(ys.shape[0]), _p(out))
    return out


def interpolate_iota_fast(ys: np.ndarray, threads: Optional[int] = None) -> np.ndarray:
    ys = _u64(ys).reshape(-1, 4)
    out = np.empty_like(ys)
    lib().orc_interpolate_iota_fast(_p(ys), C.c_size_t(ys.shape[0]), _p(out), C.c_int(threads or ncpu()))
    return out


def horner(coeffs: np.ndarray, x: np.ndarray) -> np.ndarray:
    coeffs = _u64(coeffs).reshape(-1, 4); x = _u64(x).reshape(4)
    out = np.empty(4, dtype=np.uint64)
    lib().orc_horner(_p(coeffs), C.c_size_t(coeffs.shape[0]), _p(x), _p(out))
    return out


def ntt(a: np.ndarray, inverse: bool = False, threads: int = 1) -> np.ndarray:
    a = _u64(a).reshape(-1, 4).copy()
    logn = a.shape[0].bit_length() - 1
    lib().orc_ntt(_p(a), C.c_uint(logn), C.c_int(1 if inverse else 0), C.c_int(threads))
    return a


def mle_evaluate(evals: np.ndarray, point: np.ndarray, threads: Optional[int] = None, fold: bool = False) -> np.ndarray:
    evals = _u64(evals).reshape(-1, 4); point = _u64(point).reshape(-1, 4)
    nv = evals.shape[0].bit_length() - 1
    assert point.shape[0] == nv
    out = np.empty(4, dtype=np.uint64)
    fn = lib().orc_mle_evaluate_fold if fold else lib().orc_mle_evaluate
    fn(_p(evals), C.c_uint(nv), _p(point), _p(out), C.c_int(threads or ncpu()))
    return out


def mle_partial_evaluate(evals: np.ndarray, fixed: np.ndarray, threads: Optional[int] = None, fold: bool = False) -> np.ndarray:
    evals = _u64(evals).reshape(-1, 4); fixed = _u64(fixed).reshape(-1, 4)
    nv = evals.shape[0].bit_length() - 1
    k = fixed.shape[0]
    out = np.empty((1 << (nv - k), 4), dtype=np.uint64)
    fn = lib().orc_mle_partial_evaluate_fold if fold else lib().orc_mle_partial_evaluate
    fn(_p(evals), C.c_uint(nv), _p(fixed), C.c_uint(k), _p(out), C.c_int(threads or ncpu()))
    return out


def eq_table(w: np.ndarray) -> np.ndarray:
    w = _u64(w).reshape(-1, 4)
    out = np.empty((1 << w.shape[0], 4), dtype=np.uint64)
    lib().orc_eq_table(_p(w), C.c_uint(w.shape[0]), _p(out))
    return out


def lt_table(nv: int) -> np.ndarray:
    out = np.empty((1 << (2 * nv), 4), dtype=np.uint64)
    lib().orc_lt_table(C.c_uint(nv), _p(out))
    return out


class SumCheckError(Exception):
    pass


def sumcheck_prove_product(tables: Sequence[np.ndarray], claimed: np.ndarray, transcript: Optional[Transcript] = None,
                           mode: str = "tables", threads: Optional[int] = None):
    """Returns dict(round_polynomials (nv,4,4), final_evaluation (4,), challenges (nv,4), finals (d,4))."""
    tabs = [_u64(t).reshape(-1, 4) for t in tables]
    d = len(tabs); n = tabs[0].shape[0]; nv = n.bit_length() - 1
    ptrs = (C.c_void_p * d)(*[t.ctypes.data for t in tabs])
    rp = np.zeros((nv, 4, 4), dtype=np.uint64); fe = np.zeros(4, dtype=np.uint64)
    ch = np.zeros((nv, 4), dtype=np.uint64); fin = np.zeros((d, 4), dtype=np.uint64)
    claimed = _u64(claimed).reshape(4)
    rc = lib().orc_sumcheck_prove_product(ptrs, C.c_int(d), C.c_uint(nv), _p(claimed),
                                          transcript._h if transcript else None,
                                          C.c_int(0 if mode == "closure" else 1), _p(rp), _p(fe), _p(ch), _p(fin),
                                          C.c_int(threads or ncpu()))
    if rc != 0:
        raise SumCheckError(f"Round {-rc - 1} consistency check failed")
    return dict(round_polynomials=rp, final_evaluation=fe, challenges=ch, finals=fin)


def sumcheck_verify(nv: int, claimed: np.ndarray, round_polys: np.ndarray, final_eval: np.ndarray,
                    transcript: Optional[Transcript] = None):
    rp = _u64(round_polys).reshape(-1, 4, 4)
    ch = np.zeros((max(rp.shape[0], 1), 4), dtype=np.uint64)
    rc = lib().orc_sumcheck_verify(C.c_uint(nv), _p(_u64(claimed).reshape(4)), _p(rp), C.c_uint(rp.shape[0]),
                                   _p(_u64(final_eval).reshape(4)), transcript._h if transcript else None, _p(ch))
    if rc < 0:
        raise SumCheckError("Proof has wrong number of rounds")
    return bool(rc), ch[:rp.shape[0]]


# ----------------------------------------------------------------- G1 / KZG / setup
def g1_generator() -> np.ndarray:
    out = np.empty(12, dtype=np.uint64); lib().orc_g1_generator(_p(out)); return out


def g1_add(a, b) -> np.ndarray:
    out = np.empty(12, dtype=np.uint64); lib().orc_g1_add(_p(_u64(a)), _p(_u64(b)), _p(out)); return out


def g1_mul(a, k) -> np.ndarray:
    out = np.empty(12, dtype=np.uint64); lib().orc_g1_mul(_p(_u64(a)), _p(_u64(k)), _p(out)); return out


def g1_equal(a, b) -> bool:
    return bool(lib().orc_g1_equal(_p(_u64(a)), _p(_u64(b))))


def g1_compress(a) -> bytes:
    a = _u64(a).reshape(-1, 12)
    out = np.empty(32 * a.shape[0], dtype=np.uint8)
    lib().orc_g1_compress(_p(a), C.c_size_t(a.shape[0]), _p(out))
    return out.tobytes()


def g1_hash(a) -> np.ndarray:
    out = np.empty(4, dtype=np.uint64); lib().orc_g1_hash(_p(_u64(a)), _p(out)); return out


def g1_batch_to_affine(a, threads: Optional[int] = None) -> np.ndarray:
    a = _u64(a).reshape(-1, 12)
    out = np.empty((a.shape[0], 8), dtype=np.uint64)
    lib().orc_g1_batch_to_affine(_p(a), C.c_size_t(a.shape[0]), _p(out), C.c_int(threads or ncpu()))
    return out


def g1_affine_canonical(a) -> List[Optional[Tuple[int, int]]]:
    a = _u64(a).reshape(-1, 12)
    out = np.empty(64 * a.shape[0], dtype=np.uint8)
    lib().orc_g1_to_affine_canonical(_p(a), C.c_size_t(a.shape[0]), _p(out))
    res = []
    for i in range(a.shape[0]):
        x = int.from_bytes(out[64 * i:64 * i + 32].tobytes(), "little")
        y = int.from_bytes(out[64 * i + 32:64 * i + 64].tobytes(), "little")
        res.append(None if x == 0 and y == 0 else (x, y))
    return res


def setup_scalars() -> Tuple[np.ndarray, bytes]:
    tau = np.empty(4, dtype=np.uint64); seed = C.create_string_buffer(32)
    lib().orc_setup_scalars(_p(tau), seed)
    return tau, seed.raw


def setup_g1_powers(count: int, fast: bool = True, threads: Optional[int] = None) -> np.ndarray:
    """g1_powers[0..count) of setup_params (src/utils.rs:89-96), Jacobian."""
    out = np.empty((count, 12), dtype=np.uint64)
    lib().orc_setup_g1_powers(C.c_size_t(count), _p(out), C.c_int(1 if fast else 0), C.c_int(threads or ncpu()))
    return out


class CommitmentError(Exception):
    pass


def kzg_commit(powers_jac: np.ndarray, poly: np.ndarray) -> np.ndarray:
    """verbatim commitments.rs:162-180"""
    powers_jac = _u64(powers_jac).reshape(-1, 12); poly = _u64(poly).reshape(-1, 4)
    out = np.empty(12, dtype=np.uint64)
    rc = lib().orc_kzg_commit(_p(powers_jac), C.c_size_t(powers_jac.shape[0]), _p(poly), C.c_size_t(poly.shape[0]), _p(out))
    if rc:
        raise CommitmentError("Polynomial degree exceeds setup size")
    return out


def msm_pippenger(bases_affine: np.ndarray, scalars: np.ndarray, threads: Optional[int] = None) -> np.ndarray:
    bases_affine = _u64(bases_affine).reshape(-1, 8); scalars = _u64(scalars).reshape(-1, 4)
    n = scalars.shape[0]
    assert bases_affine.shape[0] >= n
    out = np.empty(12, dtype=np.uint64)
    lib().orc_msm_pippenger(_p(bases_affine), _p(scalars), C.c_size_t(n), _p(out), C.c_int(threads or ncpu()))
    return out


def kzg_value_quotient(poly: np.ndarray, z: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    poly = _u64(poly).reshape(-1, 4)
    n = poly.shape[0]
    v = np.empty(4, dtype=np.uint64); q = np.empty((max(n - 1, 0), 4), dtype=np.uint64)
    lib().orc_kzg_value_quotient(_p(poly), C.c_size_t(n), _p(_u64(z).reshape(4)), _p(v), _p(q))
    return v, q


def kzg_check_trapdoor(Cm, z, v, pi) -> bool:
    return bool(lib().orc_kzg_check_trapdoor(_p(_u64(Cm)), _p(_u64(z)), _p(_u64(v)), _p(_u64(pi))))


_ERR = {1: "InvalidParameters", 4: "Commitment", 6: "SumCheck"}


class ProveError(Exception):
    def __init__(self, code):
        super().__init__(_ERR.get(code, str(code))); self.code = code


def twist_prove(powers_jac: np.ndarray, max_operations: int, addresses, values_mont: np.ndarray, is_write,
                fast: bool = True, threads: Optional[int] = None) -> Tuple[bytes, np.ndarray]:
    powers_jac = _u64(powers_jac).reshape(-1, 12)
    addresses = _u64(addresses).reshape(-1); n = addresses.shape[0]
    values_mont = _u64(values_mont).reshape(-1, 4)
    is_write = np.ascontiguousarray(is_write, dtype=np.uint8).reshape(-1)
    padded = 1 if n <= 1 else 1 << (n - 1).bit_length()
    cap = lib().orc_proof_max_bytes(C.c_uint(padded.bit_length() - 1))
    out = np.empty(cap, dtype=np.uint8); ln = C.c_size_t(0); z = np.zeros(4, dtype=np.uint64)
    rc = lib().orc_twist_prove(_p(powers_jac), C.c_size_t(powers_jac.shape[0]), C.c_size_t(max_operations),
                               _p(addresses), _p(values_mont), _p(is_write), C.c_size_t(n),
                               C.c_int(1 if fast else 0), C.c_int(threads or ncpu()), _p(out), C.byref(ln), _p(z))
    if rc:
        raise ProveError(rc)
    return out[:ln.value].tobytes(), z


def shout_prove(powers_jac: np.ndarray, max_operations: int, entries_mont: np.ndarray, lookup_indices,
                fast: bool = True, threads: Optional[int] = None) -> Tuple[bytes, np.ndarray]:
    powers_jac = _u64(powers_jac).reshape(-1, 12)
    entries_mont = _u64(entries_mont).reshape(-1, 4)
    lookup_indices = _u64(lookup_indices).reshape(-1); n = lookup_indices.shape[0]
    padded = 1 if n <= 1 else 1 << (n - 1).bit_length()
    cap = lib().orc_proof_max_bytes(C.c_uint(padded.bit_length() - 1))
    out = np.empty(cap, dtype=np.uint8); ln = C.c_size_t(0); z = np.zeros(4, dtype=np.uint64)
    rc = lib().orc_shout_prove(_p(powers_jac), C.c_size_t(powers_jac.shape[0]), C.c_size_t(max_operations),
                               _p(entries_mont), C.c_size_t(entries_mont.shape[0]), _p(lookup_indices), C.c_size_t(n),
                               C.c_int(1 if fast else 0), C.c_int(threads or ncpu()), _p(out), C.byref(ln), _p(z))
    if rc:
        raise ProveError(rc)
    return out[:ln.value].tobytes(), z


# ----------------------------------------------------------------- the real constraint sum-checks (non-parity mode of the product)
# The reference's Twist / Shout closures return zero (src/twist.rs:181-214, src/shout.rs:157-184).  The product offers, beside the
# byte-identical prove, the sum-checks those stubs stand for (host/read_check.cpp, host/memory_check.cpp).  Their CPU restatement is the
# reference's own SumCheck::prove (orc_sumcheck_prove_product: closure form or table form) applied to tables built here with Python integers.
def statement_digest_elements(domain: bytes, header: Sequence[int], segments: Sequence[bytes]) -> np.ndarray:
    """The binding digest both constraint sum-checks absorb before their first challenge (host/statement_digest.hpp): a two-level
    BLAKE2b-256 tree over 2^20-byte chunks, returned as the two field elements (low / high 128 bits) that enter the transcript."""
    import hashlib
    CH = 1 << 20
    root = hashlib.blake2b(digest_size=32)
    root.update(domain[:16].ljust(16, b"\0"))
    root.update(len(header).to_bytes(8, "little"))
    for h in header:
        root.update(int(h).to_bytes(8, "little"))
    root.update(len(segments).to_bytes(8, "little"))
    for seg in segments:
        root.update(len(seg).to_bytes(8, "little"))
        for off in range(0, len(seg), CH):
            root.update(hashlib.blake2b(seg[off:off + CH], digest_size=32).digest())
    d = root.digest()
    return fr_from_ints([int.from_bytes(d[:16], "little"), int.from_bytes(d[16:], "little")])


def _eq_ints(pt: np.ndarray) -> List[int]:
    return fr_to_ints(eq_table(pt.reshape(-1, 4))) if pt.shape[0] else [1]


def lt_point_ints(b_ints: Sequence[int], t: int) -> List[int]:
    """[LT~(a, b) for a in range(2^t)]: [a < c] in the natural integer order, multilinear in c, at the field point b (t canonical integers)"""
    p = R_MOD
    out = []
    for a in range(1 << t):
        prefix, acc = 1, 0
        for i in range(t - 1, -1, -1):
            bi = b_ints[i]
            if (a >> i) & 1:
                prefix = prefix * bi % p
            else:
                acc = (acc + prefix * bi) % p
                prefix = prefix * (1 - bi) % p
        out.append(acc)
    return out


def shout_read_check_prove(entries: np.ndarray, idx: np.ndarray, vals: np.ndarray, mode: str = "tables", transcript: Optional["Transcript"] = None):
    """core Shout read-checking rv~(r) = sum_x ra~(x, r) Val~(x) on a fresh transcript (or the one given: e.g. bound to a proof's commitments)
    -> (claim, sum-check result dict)"""
    p = R_MOD
    nent, nlook = entries.shape[0], idx.shape[0]
    K = 1 << max(nent - 1, 0).bit_length(); L = 1 << max(nlook - 1, 0).bit_length()
    l = L.bit_length() - 1
    tr = transcript if transcript is not None else Transcript()
    tr.append_field_elements(b"read_check_statement", statement_digest_elements(
        b"shout_read_check", [nent, nlook], [_u64(entries).tobytes(), _u64(idx).tobytes(), _u64(vals).tobytes()]))
    r = tr.challenge_field_elements(b"read_check_point", l)
    eq = _eq_ints(r)
    vi = fr_to_ints(vals) if nlook else []
    claim = sum(e * v for e, v in zip(eq, vi)) % p
    claim_fr = fr_from_ints([claim])[0]
    tr.append_field_element(b"read_check_claim", claim_fr)
    A = [0] * K
    for j in range(nlook):
        A[int(idx[j])] = (A[int(idx[j])] + eq[j]) % p
    V = fr_to_ints(entries) + [0] * (K - nent)
    return claim_fr, sumcheck_prove_product([fr_from_ints(A), fr_from_ints(V)], claim_fr, transcript=tr, mode=mode)


def twist_memory_check_prove(addr: np.ndarray, vals: np.ndarray, isw: np.ndarray, K: int, mode: str = "tables", with_write_check: bool = False,
                             transcript: Optional["Transcript"] = None):
    """Twist read-checking over (cell, cycle) + Val-evaluation on a fresh transcript -> (read claim, Val~(x*, j*), part 1, part 2); with_write_check
    appends (write claim, Val~(x**, j**), part 3, part 4): write-checking and its Val-evaluation on the same transcript"""
    p = R_MOD
    n = addr.shape[0]
    T = 1 << max(n - 1, 0).bit_length()
    k, t = K.bit_length() - 1, T.bit_length() - 1
    vi = fr_to_ints(vals) if n else []
    tr = transcript if transcript is not None else Transcript()
    tr.append_field_elements(b"memory_check_statement", statement_digest_elements(
        b"twist_memory_chk", [n, K], [_u64(addr).tobytes(), _u64(vals).tobytes(), np.ascontiguousarray(isw, dtype=np.uint8).tobytes()]))
    r = tr.challenge_field_elements(b"memory_check_point", t)
    eq = _eq_ints(r)
    claim1 = sum(eq[j] * vi[j] for j in range(n) if not isw[j]) % p
    claim1_fr = fr_from_ints([claim1])[0]
    tr.append_field_element(b"memory_read_claim", claim1_fr)
    RA = [0] * (K * T); VAL = [0] * (K * T)          # tables over index x + K j
    mem = [0] * K; inc = [0] * T
    for j in range(T):
        for x in range(K):
            VAL[x + K * j] = mem[x]
        if j < n:
            a = int(addr[j])
            if isw[j]:
                inc[j] = (vi[j] - mem[a]) % p
                mem[a] = vi[j]
            else:
                RA[a + K * j] = eq[j]
    ref1 = sumcheck_prove_product([fr_from_ints(RA), fr_from_ints(VAL)], claim1_fr, transcript=tr, mode=mode)
    ch = ref1["challenges"].reshape(-1, 4)
    val_claim = mle_evaluate(fr_from_ints(VAL), ch, fold=(mode == "tables")).reshape(4)   # Val~(x*, j*): MultilinearExtension::evaluate
    tr.append_field_element(b"memory_val_claim", val_claim)
    x_star, j_star = ch[:k], ch[k:]
    eqx = _eq_ints(x_star)
    U = [inc[j] * eqx[int(addr[j])] % p if j < n else 0 for j in range(T)]
    V = lt_point_ints(fr_to_ints(j_star) if t else [], t)
    ref2 = sumcheck_prove_product([fr_from_ints(U), fr_from_ints(V)], val_claim, transcript=tr, mode=mode)
    if not with_write_check:
        return claim1_fr, val_claim, ref1, ref2
    # ---- write-checking (host/memory_check.cpp, tsgpu_twist_write_check_prove) on the same transcript:
    #      sum_j eq(r', j) Inc_j = sum_{x, j} ( eq(r', j) [write_j] ra(x, j) ) * ( value_j - Val(x, j) ), then Val-evaluation of the Val~ claim it ends in
    tr.append_field_elements(b"memory_write_statement", statement_digest_elements(
        b"twist_memory_chk", [n, K], [_u64(addr).tobytes(), _u64(vals).tobytes(), np.ascontiguousarray(isw, dtype=np.uint8).tobytes()]))
    rw = tr.challenge_field_elements(b"memory_write_point", t)
    eqw = _eq_ints(rw)
    claim3 = sum(eqw[j] * inc[j] for j in range(T)) % p
    claim3_fr = fr_from_ints([claim3])[0]
    tr.append_field_element(b"memory_write_claim", claim3_fr)
    WA = [0] * (K * T)
    D = [0] * (K * T)
    for j in range(T):
        vj = vi[j] if j < n else 0
        for x in range(K):
            D[x + K * j] = (vj - VAL[x + K * j]) % p
        if j < n and isw[j]:
            WA[int(addr[j]) + K * j] = eqw[j]
    ref3 = sumcheck_prove_product([fr_from_ints(WA), fr_from_ints(D)], claim3_fr, transcript=tr, mode=mode)
    ch3 = ref3["challenges"].reshape(-1, 4)
    x2, j2 = ch3[:k], ch3[k:]
    wv_at = mle_evaluate(fr_from_ints([vi[j] if j < n else 0 for j in range(T)]), j2, fold=(mode == "tables")).reshape(4) if t else fr_from_ints([vi[0] if n else 0])[0]
    d_at = mle_evaluate(fr_from_ints(D), ch3, fold=(mode == "tables")).reshape(4)           # second factor at (x**, j**): MultilinearExtension::evaluate
    val_claim2 = field_binop("fr", "sub", wv_at.reshape(1, 4), d_at.reshape(1, 4))[0]
    tr.append_field_element(b"memory_val_claim_2", val_claim2)
    eqx2 = _eq_ints(x2)
    U2 = [inc[j] * eqx2[int(addr[j])] % p if j < n else 0 for j in range(T)]
    V2 = lt_point_ints(fr_to_ints(j2) if t else [], t)
    ref4 = sumcheck_prove_product([fr_from_ints(U2), fr_from_ints(V2)], val_claim2, transcript=tr, mode=mode)
    return claim1_fr, val_claim, ref1, ref2, claim3_fr, val_claim2, ref3, ref4
