"""GPU parity: vector_to_polynomial = poly_utils::lagrange_interpolate over x_i = i
(src/twist.rs:307-315, src/polynomials.rs:301-352) vs the oracle's verbatim O(n^3) restatement (small n)
and its NTT version (large n) - bit-exact."""
import numpy as np
import pytest

from conftest import seed_bytes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 32, 64, 128, 256])
def test_matches_verbatim_lagrange(ctx, oracle, n):
    vals = oracle.chacha_fr_rand(seed_bytes(n), n).reshape(n, 4)
    xs = oracle.fr_from_ints(list(range(n)))
    assert (ctx.interpolate_iota(vals) == oracle.lagrange_interpolate(xs, vals)).all()


def test_reference_anchor_parabola(ctx, oracle):
    """tests/polynomial_tests.rs:191-208: (0,0),(1,1),(2,4) -> x^2; padded with (3,9) it stays [0,0,1,0]."""
    got = ctx.interpolate_iota(oracle.fr_from_ints([0, 1, 4, 9]))
    assert oracle.fr_to_ints(got) == [0, 0, 1, 0]


@pytest.mark.parametrize("logn", [10, 13, 16])
def test_matches_fast_oracle(ctx, oracle, logn):
    n = 1 << logn
    vals = oracle.chacha_fr_rand(seed_bytes(logn), n)
    assert (ctx.interpolate_iota(vals) == oracle.interpolate_iota_fast(vals)).all()


def test_structured_vectors(ctx, oracle):
    """what the protocols actually feed in: small addresses, zero padding, constant vectors"""
    n = 1 << 11
    for name, ints in (("zeros", [0] * n), ("const", [7] * n), ("iota", list(range(n))), ("padded", list(range(1, 1000)) + [0] * (n - 999)),
                       ("squares", [i * i for i in range(n)])):
        vals = oracle.fr_from_ints(ints)
        got = ctx.interpolate_iota(vals)
        assert (got == oracle.interpolate_iota_fast(vals)).all(), name
    assert oracle.fr_to_ints(ctx.interpolate_iota(oracle.fr_from_ints([7] * 64))) == [7] + [0] * 63
    assert oracle.fr_to_ints(ctx.interpolate_iota(oracle.fr_from_ints(list(range(64))))) == [0, 1] + [0] * 62


def test_large_roundtrip_by_evaluation(ctx, oracle):
    """2^20 values: the coefficients must reproduce the values at sampled points (size-independent check),
    and equal the CPU NTT oracle."""
    n = 1 << 20
    vals = oracle.chacha_fr_rand(seed_bytes(99), n)
    got = ctx.interpolate_iota(vals)
    for i in (0, 1, 2, 12345, n // 2, n - 2, n - 1):
        assert (oracle.horner(got, oracle.fr_from_ints([i])[0]) == vals[i]).all()
    assert (got == oracle.interpolate_iota_fast(vals)).all()


def test_non_power_of_two_lengths(ctx, tsgpu, oracle):
    """the host-buffer entry point interpolates any length (reference semantics: n points, degree < n); the in-place device form,
    which Twist/Shout feed with padded vectors, still insists on a power of two"""
    got = ctx.interpolate_iota(oracle.fr_from_ints([1, 2, 3]))
    assert (got == oracle.lagrange_interpolate(oracle.fr_from_ints([0, 1, 2]), oracle.fr_from_ints([1, 2, 3]))).all()
    p = ctx.poly_upload(oracle.fr_from_ints([1, 2, 3]))
    with pytest.raises(tsgpu.TwistAndShoutError) as e:
        p.interpolate_iota()
    assert e.value.variant == "Polynomial"
