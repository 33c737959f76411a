"""GPU parity of the evaluation-basis KZG path (csrc/lagrange.cu): commit / open of vector_to_polynomial(values)
computed from the VALUES with the Lagrange-basis SRS must give the same group elements and field elements as
the reference pipeline interpolate -> commit / open on coefficients (src/twist.rs:151-160,226-243,307-315,
src/commitments.rs:162-199), and Twist/Shout proofs must have the same bytes on both paths.  Bit-exact."""
import numpy as np
import pytest

from conftest import seed_bytes

pytestmark = pytest.mark.gpu


def _oracle_interpolate(oracle, vals):
    """vector_to_polynomial on the CPU: the reference's own O(n^3) lagrange_interpolate for small n, the fast tier above"""
    n = vals.shape[0]
    if n <= 32:
        return oracle.lagrange_interpolate(oracle.fr_from_ints(list(range(n))), vals)
    return oracle.interpolate_iota_fast(vals)


@pytest.fixture(scope="module")
def srs12(ctx, oracle):
    tau, _ = oracle.setup_scalars()
    n = (1 << 12) + 1
    return ctx.srs_generate(tau, n), oracle.setup_g1_powers(n, fast=True)


@pytest.mark.parametrize("n", [1, 2, 4, 32, 64, 1024, 4096])
def test_commit_values_equals_interpolate_then_commit(ctx, tsgpu, oracle, srs12, n):
    srs, ref = srs12
    assert srs.can_lagrange()
    vals = oracle.chacha_fr_rand(seed_bytes(n + 11), n).reshape(n, 4)
    pv = ctx.poly_upload(vals)
    got = tsgpu.KZGCommitment.commit_values(srs, pv)
    assert srs.has_lagrange(n)
    coeffs = ctx.interpolate_iota(vals)
    want_dev = tsgpu.KZGCommitment.commit(srs, coeffs)
    assert tsgpu.g1_compress(got) == tsgpu.g1_compress(want_dev)
    # and against the CPU oracle: interpolation on x_i = i, then the commitment sum
    ocoeffs = _oracle_interpolate(oracle, vals)
    want = oracle.kzg_commit(ref, ocoeffs) if n <= 1024 else oracle.msm_pippenger(oracle.g1_batch_to_affine(ref[:n]), ocoeffs)
    assert tsgpu.g1_compress(got) == oracle.g1_compress(want)


@pytest.mark.parametrize("n", [1, 2, 8, 33, 500, 4096])
def test_commit_values_small_and_sparse_scalars(ctx, tsgpu, oracle, srs12, n):
    """what Twist actually commits to: 16-bit addresses, u64 values, zero padding"""
    srs, _ = srs12
    padded = 1
    while padded < n:
        padded <<= 1
    rng = np.random.default_rng(n)
    for ints in (rng.integers(0, 1 << 16, size=n), rng.integers(0, 1 << 63, size=n), np.zeros(n, dtype=np.int64), np.full(n, 65535)):
        v = ctx.poly_from_u64(np.asarray(ints, dtype=np.uint64), padded)
        got = tsgpu.KZGCommitment.commit_values(srs, v)
        c = v.clone().interpolate_iota()
        assert tsgpu.g1_compress(got) == tsgpu.g1_compress(tsgpu.KZGCommitment.commit(srs, c))


@pytest.mark.parametrize("n", [1, 2, 4, 32, 1024, 4096])
def test_open_values_equals_open_on_coefficients(ctx, tsgpu, oracle, srs12, n):
    srs, ref = srs12
    vals = oracle.chacha_fr_rand(seed_bytes(n + 21), n).reshape(n, 4)
    z = oracle.chacha_fr_rand(seed_bytes(n + 22), 1).reshape(1, 4)
    pv = ctx.poly_upload(vals)
    value, proof = tsgpu.KZGCommitment.open_values(srs, pv, z)
    coeffs = ctx.interpolate_iota(vals)
    value_c, proof_c = tsgpu.KZGCommitment.open(srs, coeffs, z)
    assert (value == value_c).all()
    assert tsgpu.g1_compress(proof) == tsgpu.g1_compress(proof_c)
    if n <= 1024:
        ov, oq = oracle.kzg_value_quotient(_oracle_interpolate(oracle, vals), z[0])
        assert (value == ov).all() and tsgpu.g1_compress(proof) == oracle.g1_compress(oracle.kzg_commit(ref, oq))


def test_open_values_at_a_node_is_refused(ctx, tsgpu, oracle, srs12):
    srs, _ = srs12
    pv = ctx.poly_upload(oracle.chacha_fr_rand(seed_bytes(5), 16).reshape(16, 4))
    with pytest.raises(tsgpu.TwistAndShoutError) as e:
        tsgpu.KZGCommitment.open_values(srs, pv, tsgpu.fe(7).reshape(1, 4))
    assert e.value.variant == "Polynomial"
    # just outside the node range is fine
    v, _ = tsgpu.KZGCommitment.open_values(srs, pv, tsgpu.fe(16).reshape(1, 4))
    c = ctx.interpolate_iota(pv.download())
    vc, _ = tsgpu.KZGCommitment.open(srs, c, tsgpu.fe(16).reshape(1, 4))
    assert (v == vc).all()


def test_uploaded_srs_has_no_evaluation_basis(ctx, tsgpu, oracle, srs12):
    _, ref = srs12
    up = ctx.srs_upload(ref[:65])
    assert not up.can_lagrange()
    pv = ctx.poly_upload(oracle.chacha_fr_rand(seed_bytes(6), 64).reshape(64, 4))
    with pytest.raises(tsgpu.TwistAndShoutError) as e:
        tsgpu.KZGCommitment.commit_values(up, pv)
    assert e.value.variant == "InvalidParameters"


@pytest.mark.parametrize("nops", [1, 2, 5, 33, 1000, 4096, 1 << 14])
def test_twist_proof_bytes_identical_on_both_paths(ctx, tsgpu, oracle, nops):
    pp, vp = tsgpu.setup_params(ctx, 12)
    rng = np.random.default_rng(nops + 1)
    addr = rng.integers(0, 1 << 12, size=nops).astype(np.uint64)
    vals = tsgpu.fe_vec([int(x) for x in rng.integers(0, 1 << 63, size=nops)])
    isw = rng.integers(0, 2, size=nops).astype(np.uint8)
    twist = tsgpu.Twist.new(pp)
    try:
        ctx.set_tuning("eval_basis", 1)
        a = twist.prove_arrays(addr, vals, isw)
        ctx.set_tuning("eval_basis", 0)
        b = twist.prove_arrays(addr, vals, isw)
    finally:
        ctx.set_tuning("eval_basis", 1)
    assert a.to_bytes() == b.to_bytes()
    assert twist.verify(a, vp)
    if nops <= 4096:
        want, _ = oracle.twist_prove(pp.srs.download(), pp.max_operations, addr, vals, isw, fast=True)
        assert a.to_bytes() == want


@pytest.mark.parametrize("nent,nlook", [(3, 2), (100, 37), (1000, 4096), (1 << 12, 1 << 14)])
def test_shout_proof_bytes_identical_on_both_paths(ctx, tsgpu, oracle, nent, nlook):
    pp, vp = tsgpu.setup_params(ctx, 12)
    rng = np.random.default_rng(nent + nlook)
    entries = tsgpu.fe_vec([i * i for i in range(nent)])
    idx = rng.integers(0, nent, size=nlook).astype(np.uint64)
    shout = tsgpu.Shout.new(pp)
    try:
        ctx.set_tuning("eval_basis", 1)
        a = shout.prove_arrays(entries, idx)
        ctx.set_tuning("eval_basis", 0)
        b = shout.prove_arrays(entries, idx)
    finally:
        ctx.set_tuning("eval_basis", 1)
    assert a.to_bytes() == b.to_bytes()
    assert shout.verify(a, vp)
    if nlook <= 4096:
        want, _ = oracle.shout_prove(pp.srs.download(), pp.max_operations, entries, idx, fast=True)
        assert a.to_bytes() == want
