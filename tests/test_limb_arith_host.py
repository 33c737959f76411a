"""CPU-only check of the exact limb algorithms the CUDA kernels run (csrc/fp.cuh), compiled for the host
with the PTX carry-flag primitives emulated (csrc/ptx.cuh), against the CPU oracle - bit-exact.
This is how the field arithmetic is validated in a container without a GPU; the GPU parity tests
(-m gpu) then check the compiled device code end to end."""
import ctypes as C
import os
import random
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# the second build runs the same checks under UBSan (shift widths, signed overflow, misaligned access abort the process): the host-side
# stand-in for compute-sanitizer, which is closed on the GPU pool
@pytest.fixture(scope="module", params=["", "-fsanitize=undefined -fno-sanitize-recover=all"], ids=["plain", "ubsan"])
def harness(request):
    out = os.path.join(tempfile.gettempdir(), f"tsg_limb_harness_{os.getpid()}_{request.param_index}.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", *request.param.split(), *os.environ.get("TSG_HARNESS_CXXFLAGS", "").split(),
                           "-I", os.path.join(ROOT, "multilinear-map-cryptography_b200", "csrc"),
                           "-o", out, os.path.join(ROOT, "tests", "helpers", "limb_harness.cpp")])
    lib = C.CDLL(out)
    yield lib
    os.unlink(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _cases(mod, n, seed):
    rnd = random.Random(seed)
    edge = [0, 1, 2, mod - 1, mod - 2, 1 << 253, (1 << 32) - 1, (1 << 64) - 1, mod >> 1, (1 << 29) - 1, 1 << 29,
            (1 << 232) - 1, 1 << 232]
    xs = edge + [rnd.randrange(mod) for _ in range(n)]
    ys = [rnd.choice(edge) for _ in edge] + [rnd.randrange(mod) for _ in range(n)]
    return xs, ys


@pytest.mark.parametrize("field", ["fr", "fq"])
def test_add_sub_mul_match_oracle(harness, oracle, field):
    mod = oracle.R_MOD if field == "fr" else oracle.P_MOD
    xs, ys = _cases(mod, 4000, 11)
    a = oracle.ints_to_limbs(xs); b = oracle.ints_to_limbs(ys); n = len(xs)
    f = 0 if field == "fr" else 1
    # op 2: product path the kernels use; op 3: carry-chain wide product + normalise + reduce; op 8: carry-chain product
    # op 10: multiplication through the per-launch constant table (sum-check fold)
    for op, name in ((0, "add"), (1, "sub"), (2, "mul"), (3, "mul"), (8, "mul"), (10, "mul")):
        out = np.empty_like(a)
        harness.limb_binop(f, op, _p(a), _p(b), C.c_size_t(n), _p(out))
        assert (out == oracle.field_binop(field, name, a, b)).all(), (field, op)


@pytest.mark.parametrize("field", ["fr", "fq"])
def test_montgomery_conversions_and_inverse(harness, oracle, field):
    mod = oracle.R_MOD if field == "fr" else oracle.P_MOD
    xs, ys = _cases(mod, 300, 5)
    a = oracle.ints_to_limbs(xs); b = oracle.ints_to_limbs(ys); n = len(xs)
    f = 0 if field == "fr" else 1
    rinv = pow(1 << 256, -1, mod)
    out = np.empty_like(a); harness.limb_binop(f, 4, _p(a), _p(b), C.c_size_t(n), _p(out))
    assert oracle.limbs_to_ints(out) == [x * rinv % mod for x in xs]
    out = np.empty_like(a); harness.limb_binop(f, 5, _p(a), _p(b), C.c_size_t(n), _p(out))
    assert oracle.limbs_to_ints(out) == [x * (1 << 256) % mod for x in xs]
    out = np.empty_like(a); harness.limb_binop(f, 7, _p(a), _p(b), C.c_size_t(n), _p(out))
    assert oracle.limbs_to_ints(out) == [(-x) % mod for x in xs]
    m = 24
    out = np.empty_like(a[:m]); harness.limb_binop(f, 6, _p(a[:m].copy()), _p(b), C.c_size_t(m), _p(out))
    R = 1 << 256
    assert oracle.limbs_to_ints(out) == [(pow(x * rinv % mod, -1, mod) * R % mod if x else 0) for x in xs[:m]]


@pytest.mark.parametrize("field", ["fr", "fq"])
def test_lazy_dot_product(harness, oracle, field):
    mod = oracle.R_MOD if field == "fr" else oracle.P_MOD
    f = 0 if field == "fr" else 1
    rinv = pow(1 << 256, -1, mod)
    xs, ys = _cases(mod, 1000, 7)
    a = oracle.ints_to_limbs(xs); b = oracle.ints_to_limbs(ys)
    w = np.zeros(8, dtype=np.uint64)
    for i in range(40):
        harness.limb_mul_wide(_p(a[i].copy()), _p(b[i].copy()), _p(w))
        assert sum(int(w[k]) << (64 * k) for k in range(8)) == xs[i] * ys[i]
    for n in (0, 1, 15, 16, 17, 100, len(xs)):
        o = np.zeros(4, dtype=np.uint64)
        for fn in (harness.limb_dot,):
            fn(f, _p(a), _p(b), C.c_size_t(n), _p(o))
            assert oracle.limbs_to_ints(o)[0] == sum(x * y for x, y in zip(xs[:n], ys[:n])) * rinv % mod
    big = oracle.ints_to_limbs([mod - 1] * 200)     # worst case for accumulator head-room
    o = np.zeros(4, dtype=np.uint64)
    for fn in (harness.limb_dot,):
        fn(f, _p(big), _p(big), C.c_size_t(200), _p(o))
        assert oracle.limbs_to_ints(o)[0] == 200 * (mod - 1) ** 2 * rinv % mod


@pytest.mark.parametrize("field", ["fr", "fq"])
def test_mul_sub_single_reduction_matches_two_products(harness, oracle, field):
    """limb::mont_mul_sub: (a b - c d) R^-1 with ONE Montgomery reduction (the Y coordinate of the mixed point addition) equals the
    difference of two Montgomery products, including every combination of the edge values (a b < c d, equal products, zeros, p - 1)"""
    mod = oracle.R_MOD if field == "fr" else oracle.P_MOD
    rnd = random.Random(5)
    edge = [0, 1, 2, mod - 1, mod - 2, mod >> 1, (1 << 253), (1 << 128) - 1]
    quads = [(a, b, c, d) for a in edge for b in edge[:4] for c in edge[:4] for d in edge] + [(7, 9, 9, 7), (mod - 1, mod - 1, mod - 1, mod - 1)]
    quads += [tuple(rnd.randrange(mod) for _ in range(4)) for _ in range(3000)]
    cols = [oracle.ints_to_limbs([q[k] for q in quads]) for k in range(4)]
    n = len(quads)
    out = np.empty_like(cols[0])
    harness.limb_mul_sub(C.c_int(0 if field == "fr" else 1), _p(cols[0]), _p(cols[1]), _p(cols[2]), _p(cols[3]), C.c_size_t(n), _p(out))
    rinv = pow(1 << 256, -1, mod)
    want = [((a * b - c * d) * rinv) % mod for a, b, c, d in quads]
    assert oracle.limbs_to_ints(out) == want


@pytest.mark.parametrize("field", ["fr", "fq"])
def test_squaring_matches_product(harness, oracle, field):
    """fp::sqr == a * a * R^-1 on edge values and all-ones patterns of every length"""
    mod = oracle.R_MOD if field == "fr" else oracle.P_MOD
    xs, _ = _cases(mod, 6000, 21)
    xs += [(1 << 256) % mod, mod - 3, 0xFFFFFFFF, 0xFFFFFFFF << 32, ((1 << 254) - 1) % mod] + [((1 << k) - 1) % mod for k in range(1, 255, 7)]
    a = oracle.ints_to_limbs(xs); out = np.empty_like(a)
    harness.limb_binop(C.c_int(0 if field == "fr" else 1), C.c_int(11), _p(a), _p(a), C.c_size_t(len(xs)), _p(out))
    rinv = pow(1 << 256, -1, mod)
    assert oracle.limbs_to_ints(out) == [x * x * rinv % mod for x in xs]
