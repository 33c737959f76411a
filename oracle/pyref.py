"""Pure-Python (bigint) restatement of the reference's Twist/Shout prover path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product (`multilinear-map-cryptography_b200/`)
may import this module; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may.

Every function cites the reference file:line it restates (paths relative to the
reference crate root).  Arithmetic that lives in third-party crates which are
NOT vendored in the reference (ark-ff 0.4.2, ark-ec 0.4.2, ark-bn254 0.4.0,
ark-serialize 0.4.2, rand_chacha 0.3.1 / rand_core 0.6.4, Rust std 1.89
DefaultHasher = SipHash-1-3) is restated from the published algorithms.

PARITY STATUS: "parity unpinned" for the ChaCha20Rng / Fp::rand / SipHash /
ark-serialize byte behaviours - the reference ships no golden vector for them
(SURVEY.md section 4, 8c).  What *is* pinned: the primitives' published
known-answer vectors (ChaCha20 zero-key keystream, SipHash-1-3 empty input,
EIP-196 BN254 vectors), the reference tests' small numeric anchors, and the
cross-check vectors of SURVEY.md Appendix C (tests/golden/appendix_c.json).

This file is deliberately slow and simple: Python ints, `%`, `pow(x,-1,p)`.
It is the independent check for the C++ oracle (oracle/oracle.cpp), which is
in turn the checker for the CUDA path.
"""
from __future__ import annotations

import struct
from typing import Callable, List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------
# BN254 constants (ark-bn254 0.4.0)
# --------------------------------------------------------------------------
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # Fr
P_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583  # Fq
MONT_R = 1 << 256
G1_GEN = (1, 2)
CURVE_B = 3

M64 = (1 << 64) - 1
M32 = (1 << 32) - 1


def fr(x: int) -> int:
    return x % R_MOD


# --------------------------------------------------------------------------
# ChaCha20Rng  (rand_chacha 0.3.1 ChaCha20Rng over rand_core 0.6.4 BlockRng)
# --------------------------------------------------------------------------
def _rotl32(v: int, c: int) -> int:
    return ((v << c) & M32) | (v >> (32 - c))


def _chacha_block(key_words: Sequence[int], counter: int, stream: int = 0) -> List[int]:
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + [
        counter & M32, (counter >> 32) & M32, stream & M32, (stream >> 32) & M32]
    x = st[:]

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & M32; x[d] = _rotl32(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & M32; x[b] = _rotl32(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & M32; x[d] = _rotl32(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & M32; x[b] = _rotl32(x[b] ^ x[c], 7)

    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(x[i] + st[i]) & M32 for i in range(16)]


class ChaCha20Rng:
    """`ChaCha20Rng::from_seed(seed)`: 20 rounds, 64-bit block counter from 0,
    stream id 0, BlockRng buffer of 4 blocks = 64 u32 words."""

    def __init__(self, seed: bytes):
        assert len(seed) == 32
        self.key = list(struct.unpack("<8I", seed))
        self.counter = 0
        self.buf: List[int] = [0] * 64
        self.index = 64  # empty

    def _generate(self):
        out: List[int] = []
        for i in range(4):
            out += _chacha_block(self.key, self.counter + i)
        self.counter += 4
        self.buf = out

    def next_u32(self) -> int:
        if self.index >= 64:
            self._generate(); self.index = 0
        v = self.buf[self.index]; self.index += 1
        return v

    def next_u64(self) -> int:
        # rand_core 0.6.4 BlockRng::next_u64
        ln = 64
        idx = self.index
        if idx < ln - 1:
            self.index += 2
            return (self.buf[idx + 1] << 32) | self.buf[idx]
        elif idx >= ln:
            self._generate(); self.index = 2
            return (self.buf[1] << 32) | self.buf[0]
        else:
            x = self.buf[ln - 1]
            self._generate(); self.index = 1
            y = self.buf[0]
            return (y << 32) | x

    def fill_bytes(self, n: int) -> bytes:
        # BlockRng::fill_bytes: consumes whole u32 words, LE bytes
        out = b""
        while len(out) < n:
            if self.index >= 64:
                self._generate(); self.index = 0
            need = n - len(out)
            words = min((need + 3) // 4, 64 - self.index)
            chunk = b"".join(struct.pack("<I", w) for w in self.buf[self.index:self.index + words])
            out += chunk[:need]
            self.index += words
        return out


def fr_rand_mont(rng: ChaCha20Rng) -> int:
    """ark-ff 0.4.2 `Fp::rand`: 4 x next_u64 (low limb first), clear the top
    2 bits, accept if < r.  Returns the raw limbs integer, which arkworks uses
    *directly as the Montgomery representation*."""
    while True:
        limbs = [rng.next_u64() for _ in range(4)]
        limbs[3] &= M64 >> 2
        v = limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | (limbs[3] << 192)
        if v < R_MOD:
            return v


_RINV_R = pow(MONT_R, -1, R_MOD)


def fr_rand(rng: ChaCha20Rng) -> int:
    """Canonical value of `Fr::rand(rng)`."""
    return fr_rand_mont(rng) * _RINV_R % R_MOD


# --------------------------------------------------------------------------
# SipHash-1-3 with keys (0,0)  == Rust std DefaultHasher::new()
# --------------------------------------------------------------------------
def _rotl64(v: int, c: int) -> int:
    return ((v << c) & M64) | (v >> (64 - c))


def siphash13(data: bytes, k0: int = 0, k1: int = 0) -> int:
    v0 = k0 ^ 0x736F6D6570736575
    v1 = k1 ^ 0x646F72616E646F6D
    v2 = k0 ^ 0x6C7967656E657261
    v3 = k1 ^ 0x7465646279746573

    def rnd():
        nonlocal v0, v1, v2, v3
        v0 = (v0 + v1) & M64; v1 = _rotl64(v1, 13); v1 ^= v0; v0 = _rotl64(v0, 32)
        v2 = (v2 + v3) & M64; v3 = _rotl64(v3, 16); v3 ^= v2
        v0 = (v0 + v3) & M64; v3 = _rotl64(v3, 21); v3 ^= v0
        v2 = (v2 + v1) & M64; v1 = _rotl64(v1, 17); v1 ^= v2; v2 = _rotl64(v2, 32)

    n = len(data)
    full = n - (n % 8)
    for i in range(0, full, 8):
        m = int.from_bytes(data[i:i + 8], "little")
        v3 ^= m; rnd(); v0 ^= m
    b = (n & 0xFF) << 56 | int.from_bytes(data[full:], "little")
    v3 ^= b; rnd(); v0 ^= b
    v2 ^= 0xFF
    rnd(); rnd(); rnd()
    return v0 ^ v1 ^ v2 ^ v3


def default_hasher_vec_u8(state: bytes) -> int:
    """`Vec<u8>::hash(&mut DefaultHasher::new()); finish()`: the slice hash
    writes the length as usize (8 bytes LE) and then the bytes."""
    return siphash13(struct.pack("<Q", len(state)) + state)


# --------------------------------------------------------------------------
# Transcript  (src/utils.rs:134-204)
# --------------------------------------------------------------------------
def fr_bytes(x: int) -> bytes:
    """ark-serialize compressed Fr/Fq: canonical integer, 32 bytes LE."""
    return (x % R_MOD).to_bytes(32, "little")


class Transcript:
    def __init__(self, seed: bytes = b"\0" * 32):
        # utils.rs:141-147 - the seeded rng is overwritten before its first use (:190)
        self.state = bytearray()

    def append_field_element(self, label: bytes, x: int):          # utils.rs:150-158
        self.state += label
        self.state += fr_bytes(x)

    def append_field_elements(self, label: bytes, xs: Sequence[int]):  # utils.rs:161-169
        self.state += label
        for x in xs:
            self.state += fr_bytes(x)

    def challenge_field_element(self, label: bytes) -> int:        # utils.rs:172-192
        self.state += label
        h = default_hasher_vec_u8(bytes(self.state))
        seed = struct.pack("<Q", h) * 4
        return fr_rand(ChaCha20Rng(seed))

    def challenge_field_elements(self, label: bytes, count: int) -> List[int]:  # utils.rs:195-203
        return [self.challenge_field_element(label + b"_" + str(i).encode()) for i in range(count)]


# --------------------------------------------------------------------------
# field utils (src/utils.rs:207-269) and poly utils (src/polynomials.rs:296-371)
# --------------------------------------------------------------------------
def horner_eval(coeffs: Sequence[int], x: int) -> int:             # utils.rs:217-221
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R_MOD
    return acc


def lagrange_interpolate(points: Sequence[Tuple[int, int]]) -> List[int]:
    """Verbatim O(n^3) loop structure of polynomials.rs:301-352."""
    n = len(points)
    if n == 0:
        return []
    result = [0] * n
    for i in range(n):
        xi, yi = points[i]
        li = [1]
        for j in range(n):
            if i == j:
                continue
            xj = points[j][0]
            dinv = pow((xi - xj) % R_MOD, -1, R_MOD)
            new = [0] * (len(li) + 1)
            for k in range(len(li)):
                new[k + 1] = (new[k + 1] + li[k]) % R_MOD
            for k in range(len(li)):
                new[k] = (new[k] - li[k] * xj) % R_MOD
            li = [c * dinv % R_MOD for c in new]
        for k in range(min(len(li), n)):
            result[k] = (result[k] + yi * li[k]) % R_MOD
    return result


def interpolate_iota(values: Sequence[int]) -> List[int]:
    """Same interpolant on x_i = i (what twist.rs:307-315 / shout.rs:277-285
    ask for), computed in O(n^2) via Newton forward differences; the interpolant
    is unique so the coefficients equal `lagrange_interpolate`'s."""
    n = len(values)
    if n == 0:
        return []
    # divided differences on 0..n-1
    d = [v % R_MOD for v in values]
    inv = [0] + [pow(k, -1, R_MOD) for k in range(1, n)]
    for k in range(1, n):
        for i in range(n - 1, k - 1, -1):
            d[i] = (d[i] - d[i - 1]) * inv[k] % R_MOD
    # Horner in Newton basis -> monomial
    coeffs = [0] * n
    for k in range(n - 1, -1, -1):
        # coeffs = coeffs * (x - k) + d[k]
        new = [0] * n
        for t in range(n - 1):
            new[t + 1] = coeffs[t]
        for t in range(n):
            new[t] = (new[t] - coeffs[t] * k) % R_MOD
        new[0] = (new[0] + d[k]) % R_MOD
        coeffs = new
    return coeffs


# --------------------------------------------------------------------------
# MultilinearExtension (src/polynomials.rs:18-196)
# --------------------------------------------------------------------------
class MultilinearExtension:
    def __init__(self, num_vars: int, evaluations: List[int]):
        self.num_vars = num_vars
        self.evaluations = evaluations

    @staticmethod
    def from_evaluations(evals: Sequence[int]) -> "MultilinearExtension":   # :28-37
        n = len(evals)
        nv = n.bit_length() - 1 if n else 0
        assert n and (1 << nv) == n, "Evaluation vector length must be a power of 2"
        return MultilinearExtension(nv, [e % R_MOD for e in evals])

    @staticmethod
    def from_evaluations_vec(num_vars: int, evals: Sequence[int]) -> "MultilinearExtension":  # :40-50
        size = 1 << num_vars
        ev = [e % R_MOD for e in evals][:size]
        ev += [0] * (size - len(ev))
        return MultilinearExtension(num_vars, ev)

    @staticmethod
    def from_sparse(num_vars: int, entries: Sequence[Tuple[int, int]]) -> "MultilinearExtension":  # :54-67
        size = 1 << num_vars
        ev = [0] * size
        for idx, v in entries:
            assert idx < size
            ev[idx] = v % R_MOD
        return MultilinearExtension(num_vars, ev)

    @staticmethod
    def one_hot(num_vars: int, index: int) -> "MultilinearExtension":       # :71-82
        size = 1 << num_vars
        assert index < size
        ev = [0] * size
        ev[index] = 1
        return MultilinearExtension(num_vars, ev)

    def evaluate(self, point: Sequence[int]) -> int:                        # :85-122
        assert len(point) == self.num_vars
        total = 0
        for index, e in enumerate(self.evaluations):
            if e == 0:
                continue
            basis = 1
            for j in range(self.num_vars):
                basis = basis * (point[j] if (index >> j) & 1 else (1 - point[j])) % R_MOD
            total = (total + e * basis) % R_MOD
        return total

    def partial_evaluate(self, fixed: Sequence[int]) -> "MultilinearExtension":  # :126-161
        k = len(fixed)
        assert k <= self.num_vars
        if k == 0:
            return MultilinearExtension(self.num_vars, list(self.evaluations))
        nn = self.num_vars - k
        out = []
        for new_index in range(1 << nn):
            full = list(fixed) + [(new_index >> j) & 1 for j in range(nn)]
            out.append(self.evaluate(full))
        return MultilinearExtension(nn, out)

    def sum_evaluations(self) -> int:                                       # :193-195
        return sum(self.evaluations) % R_MOD


def lt_bits(a: int, b: int, num_vars: int) -> int:
    """polynomials.rs:222-239: first differing bit from bit 0 upward decides."""
    for i in range(num_vars):
        ab, bb = (a >> i) & 1, (b >> i) & 1
        if ab and not bb:
            return 0
        if bb and not ab:
            return 1
    return 0


def lt_table(num_vars: int) -> List[int]:
    """polynomials.rs:243-263: index = a | (b << n)."""
    size = 1 << (2 * num_vars)
    mask = (1 << num_vars) - 1
    return [lt_bits(i & mask, i >> num_vars, num_vars) for i in range(size)]


def eq_table(w: Sequence[int]) -> List[int]:
    """eq(w, i) = prod_j (bit_j(i) ? w_j : 1-w_j) - the basis polynomial of
    polynomials.rs:108-122 tabulated over all i."""
    t = [1]
    for j, wj in enumerate(w):
        t = [x * (1 - wj) % R_MOD for x in t] + [x * wj % R_MOD for x in t]
    return t


# --------------------------------------------------------------------------
# SumCheck (src/sumcheck.rs)
# --------------------------------------------------------------------------
class SumCheckError(Exception):
    pass


def _round_coeffs_from_evals(evals: Sequence[int]) -> List[int]:
    # sumcheck.rs:201-205: lagrange_interpolate over x = 0..3
    return lagrange_interpolate([(i, evals[i]) for i in range(len(evals))])


def sumcheck_prove(num_vars: int, claimed_sum: int, f: Callable[[List[int]], int],
                   transcript: Transcript) -> Tuple[List[List[int]], int]:
    """Closure-driven sumcheck.rs:56-110 + :156-207, verbatim."""
    round_polys = []
    current = claimed_sum % R_MOD
    fixed: List[int] = []
    for rnd in range(num_vars):
        remaining = num_vars - len(fixed) - 1
        evals = []
        for x in range(4):
            s = 0
            for suffix in range(1 << remaining):
                pt = fixed + [x] + [(suffix >> b) & 1 for b in range(remaining)]
                s = (s + f(pt)) % R_MOD
            evals.append(s)
        coeffs = _round_coeffs_from_evals(evals)
        if (horner_eval(coeffs, 0) + horner_eval(coeffs, 1)) % R_MOD != current:
            raise SumCheckError(f"Round {rnd} consistency check failed")
        round_polys.append(coeffs)
        transcript.append_field_elements(f"sumcheck_round_{rnd}".encode(), coeffs)
        r = transcript.challenge_field_element(f"sumcheck_challenge_{rnd}".encode())
        fixed.append(r)
        current = horner_eval(coeffs, r)
    return round_polys, f(fixed) % R_MOD


def sumcheck_verify(num_vars: int, claimed_sum: int, round_polys: Sequence[Sequence[int]],
                    final_evaluation: int, transcript: Transcript) -> Tuple[bool, List[int]]:
    """sumcheck.rs:113-153."""
    if len(round_polys) != num_vars:
        raise SumCheckError("Proof has wrong number of rounds")
    current = claimed_sum % R_MOD
    challenges: List[int] = []
    for rnd, coeffs in enumerate(round_polys):
        if (horner_eval(coeffs, 0) + horner_eval(coeffs, 1)) % R_MOD != current:
            return False, challenges
        transcript.append_field_elements(f"sumcheck_round_{rnd}".encode(), coeffs)
        r = transcript.challenge_field_element(f"sumcheck_challenge_{rnd}".encode())
        challenges.append(r)
        current = horner_eval(coeffs, r)
    return current == final_evaluation % R_MOD, challenges


def sumcheck_prove_product_tables(tables: Sequence[Sequence[int]], claimed_sum: int,
                                  transcript: Transcript) -> Tuple[List[List[int]], int, List[int]]:
    """Linear-time table form of SumCheck::prove for f(v) = prod_t MLE_t(v):
    round k pairs entries (2i, 2i+1) (variable k <-> index bit k,
    polynomials.rs:111-118), then binds T'[i] = T[2i] + r (T[2i+1]-T[2i]).
    Returns (round_polys, final_evaluation, per-table final values)."""
    tabs = [[x % R_MOD for x in t] for t in tables]
    n = len(tabs[0])
    num_vars = n.bit_length() - 1
    current = claimed_sum % R_MOD
    round_polys = []
    for rnd in range(num_vars):
        half = len(tabs[0]) // 2
        evals = []
        for x in range(4):
            s = 0
            for i in range(half):
                prod = 1
                for t in tabs:
                    prod = prod * (t[2 * i] + x * (t[2 * i + 1] - t[2 * i])) % R_MOD
                s = (s + prod) % R_MOD
            evals.append(s)
        coeffs = _round_coeffs_from_evals(evals)
        if (horner_eval(coeffs, 0) + horner_eval(coeffs, 1)) % R_MOD != current:
            raise SumCheckError(f"Round {rnd} consistency check failed")
        round_polys.append(coeffs)
        transcript.append_field_elements(f"sumcheck_round_{rnd}".encode(), coeffs)
        r = transcript.challenge_field_element(f"sumcheck_challenge_{rnd}".encode())
        current = horner_eval(coeffs, r)
        tabs = [[(t[2 * i] + r * (t[2 * i + 1] - t[2 * i])) % R_MOD for i in range(half)] for t in tabs]
    finals = [t[0] for t in tabs]
    fe = 1
    for v in finals:
        fe = fe * v % R_MOD
    return round_polys, fe, finals


# --------------------------------------------------------------------------
# G1 (ark-ec 0.4.2 short Weierstrass, y^2 = x^3 + 3); points are None (identity)
# or affine (x, y) over Python ints.  Jacobian is used internally for speed.
# --------------------------------------------------------------------------
def _jac_double(P):
    X, Y, Z = P
    if Z == 0:
        return P
    A = X * X % P_MOD; B = Y * Y % P_MOD; C = B * B % P_MOD
    D = 2 * ((X + B) ** 2 - A - C) % P_MOD
    E = 3 * A % P_MOD; F = E * E % P_MOD
    X3 = (F - 2 * D) % P_MOD
    Y3 = (E * (D - X3) - 8 * C) % P_MOD
    Z3 = 2 * Y * Z % P_MOD
    return (X3, Y3, Z3)


def _jac_add(P, Q):
    X1, Y1, Z1 = P; X2, Y2, Z2 = Q
    if Z1 == 0:
        return Q
    if Z2 == 0:
        return P
    Z1Z1 = Z1 * Z1 % P_MOD; Z2Z2 = Z2 * Z2 % P_MOD
    U1 = X1 * Z2Z2 % P_MOD; U2 = X2 * Z1Z1 % P_MOD
    S1 = Y1 * Z2 * Z2Z2 % P_MOD; S2 = Y2 * Z1 * Z1Z1 % P_MOD
    if U1 == U2:
        if S1 == S2:
            return _jac_double(P)
        return (1, 1, 0)
    H = (U2 - U1) % P_MOD; Rr = (S2 - S1) % P_MOD
    HH = H * H % P_MOD; HHH = H * HH % P_MOD; V = U1 * HH % P_MOD
    X3 = (Rr * Rr - HHH - 2 * V) % P_MOD
    Y3 = (Rr * (V - X3) - S1 * HHH) % P_MOD
    Z3 = Z1 * Z2 * H % P_MOD
    return (X3, Y3, Z3)


def _to_jac(P):
    return (1, 1, 0) if P is None else (P[0], P[1], 1)


def _to_affine(J):
    X, Y, Z = J
    if Z == 0:
        return None
    zi = pow(Z, -1, P_MOD); zi2 = zi * zi % P_MOD
    return (X * zi2 % P_MOD, Y * zi2 * zi % P_MOD)


def g1_add(P, Q):
    return _to_affine(_jac_add(_to_jac(P), _to_jac(Q)))


def g1_neg(P):
    return None if P is None else (P[0], (-P[1]) % P_MOD)


def g1_mul(P, k: int):
    """`point * scalar`: MSB-first double-and-add over the canonical scalar."""
    k %= R_MOD
    acc = (1, 1, 0)
    J = _to_jac(P)
    for bit in bin(k)[2:] if k else "":
        acc = _jac_double(acc)
        if bit == "1":
            acc = _jac_add(acc, J)
    return _to_affine(acc)


def g1_is_on_curve(P) -> bool:
    return P is None or (P[1] * P[1] - P[0] ** 3 - CURVE_B) % P_MOD == 0


def g1_compressed(P) -> bytes:
    """ark-serialize 0.4.2 compressed SW point: x (32 B LE) with flags in the top
    two bits of the last byte: 0x80 if y > -y, 0x40 for infinity (x = 0)."""
    if P is None:
        b = bytearray(32); b[31] |= 0x40
        return bytes(b)
    x, y = P
    b = bytearray(x.to_bytes(32, "little"))
    if y > (P_MOD - y) % P_MOD:
        b[31] |= 0x80
    return bytes(b)


def g1_hash(P) -> int:
    """KZGCommitmentValue::hash, commitments.rs:73-84: affine x as integer mod r;
    identity's affine x is 0."""
    return 0 if P is None else P[0] % R_MOD


# --------------------------------------------------------------------------
# setup_params (src/utils.rs:79-131)
# --------------------------------------------------------------------------
class Params:
    def __init__(self, log_size: int, tau: int, g1_powers: list, seed: bytes):
        self.log_size = log_size
        self.max_operations = 1 << (log_size + 2)
        self.tau = tau
        self.g1_powers = g1_powers
        self.fiat_shamir_seed = seed


def setup_tau_and_seed() -> Tuple[int, int, bytes]:
    rng = ChaCha20Rng(bytes([42]) * 32)
    raw = fr_rand_mont(rng)
    tau = raw * _RINV_R % R_MOD
    seed = rng.fill_bytes(32)
    return tau, raw, seed


def setup_params(log_size: int, max_powers: Optional[int] = None) -> Params:
    tau, _, seed = setup_tau_and_seed()
    max_ops = 1 << (log_size + 2)
    max_degree = max_ops  # already a power of two (utils.rs:89)
    count = max_degree + 1 if max_powers is None else min(max_powers, max_degree + 1)
    powers = []
    cur = 1
    for _ in range(count):
        powers.append(g1_mul(G1_GEN, cur))
        cur = cur * tau % R_MOD
    return Params(log_size, tau, powers, seed)


# --------------------------------------------------------------------------
# KZG (src/commitments.rs:156-199, 305-375)
# --------------------------------------------------------------------------
class CommitmentError(Exception):
    pass


def kzg_commit(params: Params, poly: Sequence[int]):
    if len(poly) > len(params.g1_powers):
        raise CommitmentError("Polynomial degree exceeds setup size")
    acc = (1, 1, 0)
    for c, g in zip(poly, params.g1_powers):
        acc = _jac_add(acc, _to_jac(g1_mul(g, c)))
    return _to_affine(acc)


def polynomial_division_linear(poly: Sequence[int], z: int, value: int) -> List[int]:
    """compute_quotient_polynomial + polynomial_division by (x - z),
    commitments.rs:317-375 (long division from the top coefficient)."""
    if len(poly) == 0:
        return []
    rem = [c % R_MOD for c in poly]
    rem[0] = (rem[0] - value) % R_MOD
    if len(rem) < 2:
        return []
    qdeg = len(rem) - 2
    q = [0] * (qdeg + 1)
    for i in range(qdeg, -1, -1):
        coeff = rem[i + 1]
        q[i] = coeff
        rem[i] = (rem[i] + coeff * z) % R_MOD     # subtract coeff * (-z)
        rem[i + 1] = 0
    return q


def kzg_open(params: Params, poly: Sequence[int], z: int):
    value = horner_eval(poly, z) if len(poly) else 0
    q = polynomial_division_linear(poly, z, value)
    return value, kzg_commit(params, q)


def kzg_check_with_trapdoor(params: Params, C, z: int, v: int, proof) -> bool:
    """Opening identity with the retained trapdoor (utils.rs:107):
    C - v G == (tau - z) * proof.  Stands in for the pairing check."""
    lhs = g1_add(C, g1_neg(g1_mul(G1_GEN, v)))
    rhs = g1_mul(proof, (params.tau - z) % R_MOD)
    return lhs == rhs


# --------------------------------------------------------------------------
# Twist / Shout prove (src/twist.rs:107-252, src/shout.rs:97-222)
# --------------------------------------------------------------------------
def _next_pow2(n: int) -> int:
    return 1 if n <= 1 else 1 << (n - 1).bit_length()


class Proof:
    def __init__(self, c0, c1, round_polys, final_eval, openings, finals, challenges=None, z=None):
        self.commitments = (c0, c1)
        self.round_polynomials = round_polys
        self.final_evaluation = final_eval
        self.opening_proofs = openings
        self.final_evaluations = finals
        self.sumcheck_challenges = challenges
        self.z = z

    def to_bytes(self) -> bytes:
        """Canonical proof bytes, SURVEY.md Appendix D."""
        out = g1_compressed(self.commitments[0]) + g1_compressed(self.commitments[1])
        out += struct.pack("<Q", len(self.round_polynomials))
        for rp in self.round_polynomials:
            out += struct.pack("<Q", len(rp))
            for c in rp:
                out += fr_bytes(c)
        out += fr_bytes(self.final_evaluation)
        out += struct.pack("<Q", len(self.opening_proofs))
        for p in self.opening_proofs:
            out += g1_compressed(p)
        out += struct.pack("<Q", len(self.final_evaluations))
        for v in self.final_evaluations:
            out += fr_bytes(v)
        return out


def _vector_to_polynomial(vec: Sequence[int], fast: bool) -> List[int]:
    if fast:
        return interpolate_iota(vec)
    return lagrange_interpolate([(i, v) for i, v in enumerate(vec)])


def twist_prove(params: Params, ops: Sequence[Tuple[str, int, int]], fast: bool = False) -> Proof:
    """ops: ('R'|'W', address, value).  twist.rs:107-252."""
    if len(ops) > params.max_operations:
        raise ValueError("Too many operations")
    addresses = [a % R_MOD for _, a, _ in ops]
    values = [v % R_MOD for _, _, v in ops]
    op_types = [1 if k == "W" else 0 for k, _, _ in ops]
    padded = max(_next_pow2(len(ops)), 1)
    addresses += [0] * (padded - len(addresses))
    values += [0] * (padded - len(values))
    op_types += [0] * (padded - len(op_types))
    a_poly = _vector_to_polynomial(addresses, fast)
    v_poly = _vector_to_polynomial(values, fast)
    Ca = kzg_commit(params, a_poly)
    Cv = kzg_commit(params, v_poly)
    log_ops = padded.bit_length() - 1
    tr = Transcript(params.fiat_shamir_seed)
    tr.append_field_element(b"address_commitment", g1_hash(Ca))
    tr.append_field_element(b"value_commitment", g1_hash(Cv))
    a_mle = MultilinearExtension.from_evaluations_vec(log_ops, addresses)
    v_mle = MultilinearExtension.from_evaluations_vec(log_ops, values)
    o_mle = MultilinearExtension.from_evaluations_vec(log_ops, op_types)

    def closure(vs: List[int]) -> int:                      # twist.rs:191-213: zero on every branch
        if len(vs) != log_ops:
            return 0
        if not fast:
            a_mle.evaluate(vs); v_mle.evaluate(vs)
            if o_mle.evaluate(vs) == 1:
                return 0
        return 0

    rps, fe = sumcheck_prove(log_ops, 0, closure, tr)
    chal = tr.challenge_field_elements(b"opening_challenges", log_ops)
    openings, finals = [], []
    z = None
    if chal:
        z = chal[0]
        ea, pa = kzg_open(params, a_poly, z)
        ev, pv = kzg_open(params, v_poly, z)
        openings = [pa, pv]; finals = [ea, ev]
    pr = Proof(Ca, Cv, rps, fe, openings, finals, z=z)
    pr.polys = (a_poly, v_poly)
    return pr


def shout_prove(params: Params, entries: Sequence[int], lookups: Sequence[int], fast: bool = False) -> Proof:
    """entries: table; lookups: indices.  shout.rs:97-222."""
    if len(lookups) > params.max_operations:
        raise ValueError("Too many lookup operations")
    tsize = _next_pow2(len(entries))
    table = [e % R_MOD for e in entries] + [0] * (tsize - len(entries))
    lsize = max(_next_pow2(len(lookups)), 1)
    idx = [i % R_MOD for i in lookups] + [0] * (lsize - len(lookups))
    t_poly = _vector_to_polynomial(table, fast)
    i_poly = _vector_to_polynomial(idx, fast)
    Ct = kzg_commit(params, t_poly)
    Ci = kzg_commit(params, i_poly)
    log_l = lsize.bit_length() - 1
    tr = Transcript(params.fiat_shamir_seed)
    tr.append_field_element(b"table_commitment", g1_hash(Ct))
    tr.append_field_element(b"index_commitment", g1_hash(Ci))
    i_mle = MultilinearExtension.from_evaluations_vec(log_l, idx)

    def closure(vs: List[int]) -> int:                      # shout.rs:166-183
        if len(vs) != log_l:
            return 0
        if not fast:
            i_mle.evaluate(vs)
        return 0

    rps, fe = sumcheck_prove(log_l, 0, closure, tr)
    chal = tr.challenge_field_elements(b"opening_challenges", log_l)
    openings, finals = [], []
    z = None
    if chal:
        z = chal[0]
        et, pt = kzg_open(params, t_poly, z)
        ei, pi = kzg_open(params, i_poly, z)
        openings = [pt, pi]; finals = [et, ei]
    pr = Proof(Ct, Ci, rps, fe, openings, finals, z=z)
    pr.polys = (t_poly, i_poly)
    return pr
