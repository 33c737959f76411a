"""GPU parity: KZGVectorCommitment (src/commitments.rs:407-483) and interpolation on x_i = i for lengths that are NOT powers of two
(poly_utils::lagrange_interpolate, src/polynomials.rs:301-352) against the CPU oracle's verbatim restatement.  Bit-exact."""
import numpy as np
import pytest

from conftest import seed_bytes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 3, 5, 7, 31, 33, 48, 100])
def test_interpolate_any_length_matches_reference_lagrange(ctx, oracle, n):
    vals = oracle.chacha_fr_rand(seed_bytes(n + 50), n).reshape(n, 4)
    got = ctx.interpolate_iota(vals)
    want = oracle.lagrange_interpolate(oracle.fr_from_ints(list(range(n))), vals)
    assert got.shape == (n, 4) and (got == want).all()


@pytest.mark.parametrize("n", [1000, 3000, 5000])
def test_interpolate_any_length_evaluates_back(ctx, oracle, n):
    """larger ragged lengths: Horner evaluation of the coefficients at every fourth node returns the values"""
    vals = oracle.chacha_fr_rand(seed_bytes(n % 200), n).reshape(n, 4)
    coeffs = ctx.interpolate_iota(vals)
    for j in list(range(0, n, max(1, n // 7))) + [n - 1]:
        assert (oracle.horner(coeffs, oracle.fr_from_ints([j])[0]) == vals[j]).all()


@pytest.fixture(scope="module")
def params5(ctx, tsgpu):
    return tsgpu.setup_params(ctx, 5)            # 129 powers


@pytest.mark.parametrize("n", [1, 3, 8, 13, 64, 100])
def test_vector_commit_open_verify(ctx, tsgpu, oracle, params5, n):
    """commit = KZG commitment of the interpolant; open(i) proves vector[i]; verify accepts it and rejects another value"""
    pp, vp = params5
    pw = pp.srs.download()
    vec = oracle.chacha_fr_rand(seed_bytes(n + 9), n).reshape(n, 4)
    V = tsgpu.KZGVectorCommitment
    com = V.commit(pp.srs, vec)
    poly = oracle.lagrange_interpolate(oracle.fr_from_ints(list(range(n))), vec)
    assert tsgpu.g1_compress(com) == oracle.g1_compress(oracle.kzg_commit(pw, poly))
    for i in sorted({0, n // 2, n - 1}):
        value, proof = V.open(pp.srs, vec, i)
        assert (value == vec[i]).all()
        ov, oq = oracle.kzg_value_quotient(poly, oracle.fr_from_ints([i])[0])
        assert (ov == vec[i]).all() and tsgpu.g1_compress(proof) == oracle.g1_compress(oracle.kzg_commit(pw, oq))
        assert V.verify(vp, com, i, value, proof)
        assert not V.verify(vp, com, i, tsgpu.fe(12345), proof)
        if n > 1:
            assert not V.verify(vp, com, (i + 1) % n, value, proof)


def test_vector_commitment_edges(ctx, tsgpu, params5):
    pp, vp = params5
    V = tsgpu.KZGVectorCommitment
    empty = np.empty((0, 4), dtype=np.uint64)
    assert tsgpu.g1_compress(V.commit(pp.srs, empty)).hex() == "00" * 31 + "40"        # identity commitment
    vec = tsgpu.fe_vec([5, 6, 7])
    with pytest.raises(tsgpu.TwistAndShoutError) as e:
        V.open(pp.srs, vec, 3)
    assert e.value.variant == "Commitment" and "Index out of bounds" in str(e.value)
    with pytest.raises(tsgpu.TwistAndShoutError) as e:
        V.commit(pp.srs, tsgpu.fe_vec(list(range(200))))                               # longer than the 129-power SRS
    assert e.value.variant == "Commitment" and "Polynomial degree exceeds setup size" in str(e.value)
