"""GPU parity: SRS generation, G1 MSM, KZGCommitment::commit / open (src/commitments.rs:156-199,
src/utils.rs:89-96) through the C ABI vs the CPU oracle.  Group elements are compared as ark-serialize
compressed bytes (i.e. after affine normalisation) - bit-exact."""
import numpy as np
import pytest

from conftest import seed_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def srs_small(ctx, oracle):
    tau, _ = oracle.setup_scalars()
    n = (1 << 12) + 1
    srs = ctx.srs_generate(tau, n)
    ref = oracle.setup_g1_powers(n, fast=True)
    return srs, ref


def test_srs_generate_matches_setup_params(ctx, oracle, srs_small):
    """g1_powers[i] = G * tau^i (utils.rs:89-96) incl. Appendix C.1 vectors for i = 1, 32."""
    srs, ref = srs_small
    got = srs.download()
    assert oracle.g1_compress(got) == oracle.g1_compress(ref)
    assert oracle.g1_compress(got[1]).hex() == "f82610fe9c43824626b034bd432a3a7335eea949272763d214789731a135deaa"
    assert oracle.g1_compress(got[32]).hex() == "ace48dfeab869c1bbb618f4ea03e8910e1f917df382545a6cfc012da97bcbaac"
    verb = oracle.setup_g1_powers(40, fast=False)      # the reference's own serial loop
    assert oracle.g1_compress(got[:40]) == oracle.g1_compress(verb)


def test_srs_upload_roundtrip(ctx, oracle, srs_small):
    _, ref = srs_small
    # non-trivial Jacobian representatives (z != 1) and an identity entry
    pts = ref[:300].copy()
    g = oracle.g1_generator()
    pts[7] = oracle.g1_add(oracle.g1_add(g, g), g)          # z != 1
    pts[9] = oracle.g1_add(pts[3], oracle.g1_mul(pts[3], oracle.fr_from_ints([oracle.R_MOD - 1])[0]))   # identity
    srs = ctx.srs_upload(pts)
    assert oracle.g1_compress(srs.download()) == oracle.g1_compress(pts)


def test_msm_reproduces_the_published_alt_bn128_product(ctx, tsgpu, oracle):
    """EIP-196 scalar-multiplication vector "chfast1" (tests/test_published_kats.py) through the device MSM: an SRS made of copies of the published point,
    scalars that sum to the published scalar -> the published product"""
    from test_published_kats import CHFAST1_POINT, CHFAST1_SCALAR, CHFAST1_PRODUCT
    P = np.concatenate([oracle.fq_from_ints([CHFAST1_POINT[0]]).reshape(-1), oracle.fq_from_ints([CHFAST1_POINT[1]]).reshape(-1), oracle.fq_from_ints([1]).reshape(-1)])
    for n in (1, 3, 700):
        srs = ctx.srs_upload(np.stack([P] * n))
        ks = [CHFAST1_SCALAR] if n == 1 else [CHFAST1_SCALAR - sum(range(2, n + 1))] + list(range(2, n + 1))
        got = tsgpu.KZGCommitment.commit(srs, oracle.fr_from_ints(ks))
        assert oracle.g1_affine_canonical(got) == [CHFAST1_PRODUCT]


@pytest.mark.parametrize("n", [0, 1, 2, 3, 31, 32, 100, 1024])
def test_commit_matches_reference_serial_sum(ctx, tsgpu, oracle, srs_small, n):
    """vs the verbatim commitments.rs:173-177 sum of double-and-add products"""
    srs, ref = srs_small
    poly = oracle.chacha_fr_rand(seed_bytes(n + 3), n).reshape(n, 4)
    got = tsgpu.KZGCommitment.commit(srs, poly)
    want = oracle.kzg_commit(ref, poly)
    assert oracle.g1_compress(got) == oracle.g1_compress(want)
    assert tsgpu.g1_compress(got) == oracle.g1_compress(want)
    assert (tsgpu.g1_hash(got) == oracle.g1_hash(want)).all()


def test_commit_edge_scalars(ctx, tsgpu, oracle, srs_small):
    """zeros (identity commitment, hash 0), all-equal scalars (one over-full bucket per window), r-1, small values"""
    srs, ref = srs_small
    n = 3000
    for name, ints in (("zeros", [0] * n), ("ones", [1] * n), ("equal", [123456789] * n), ("max", [oracle.R_MOD - 1] * n),
                       ("small", list(range(n))), ("pow2", [1 << (i % 254) for i in range(n)])):
        poly = oracle.fr_from_ints(ints)
        got = tsgpu.KZGCommitment.commit(srs, poly)
        want = oracle.msm_pippenger(oracle.g1_batch_to_affine(ref[:n]), poly)
        assert oracle.g1_compress(got) == oracle.g1_compress(want), name
    z = tsgpu.KZGCommitment.commit(srs, oracle.fr_from_ints([0] * 8))
    assert tsgpu.g1_compress(z).hex() == "00" * 31 + "40" and oracle.fr_to_ints(tsgpu.g1_hash(z)) == [0]


def test_commit_too_long_is_commitment_error(ctx, tsgpu, oracle, srs_small):
    srs, _ = srs_small
    poly = oracle.fr_from_ints([1] * (len(srs) + 1))
    with pytest.raises(tsgpu.TwistAndShoutError) as e:
        tsgpu.KZGCommitment.commit(srs, poly)
    assert e.value.variant == "Commitment" and "Polynomial degree exceeds setup size" in str(e.value)


def test_msm_with_identity_bases_and_large_n(ctx, oracle, srs_small):
    _, ref = srs_small
    n = 4096
    aff = oracle.g1_batch_to_affine(ref[:n])
    aff[5] = 0; aff[77] = 0                                    # identity bases are skipped
    sc = oracle.chacha_fr_rand(seed_bytes(9), n)
    assert oracle.g1_compress(ctx.msm_g1(aff, sc)) == oracle.g1_compress(oracle.msm_pippenger(aff, sc))


@pytest.mark.parametrize("n", [1, 2, 5, 16, 17, 4096, 4097])
def test_open_matches_reference(ctx, tsgpu, oracle, srs_small, n):
    """value = Horner (commitments.rs:305-313), proof = commit(quotient) (:317-375, :194); trapdoor identity holds"""
    srs, ref = srs_small
    poly = oracle.chacha_fr_rand(seed_bytes(50 + n), n).reshape(n, 4)
    z = oracle.chacha_fr_rand(seed_bytes(51), 1)[0]
    value, proof = tsgpu.KZGCommitment.open(srs, poly, z)
    v_ref, q_ref = oracle.kzg_value_quotient(poly, z)
    assert (value == v_ref).all()
    if n <= 17:
        want = oracle.kzg_commit(ref, q_ref)
    else:
        want = oracle.msm_pippenger(oracle.g1_batch_to_affine(ref[:n - 1]), q_ref)
    assert oracle.g1_compress(proof) == oracle.g1_compress(want)
    C = tsgpu.KZGCommitment.commit(srs, poly)
    assert oracle.kzg_check_trapdoor(C, z, value, proof)


def test_open_reference_anchor_f5_is_86(ctx, tsgpu, oracle, srs_small):
    """src/commitments.rs:495-520: f(x) = 1 + 2x + 3x^2, f(5) = 86"""
    srs, _ = srs_small
    value, proof = tsgpu.KZGCommitment.open(srs, oracle.fr_from_ints([1, 2, 3]), oracle.fr_from_ints([5])[0])
    assert oracle.fr_to_ints(value) == [86]
    v0, p0 = tsgpu.KZGCommitment.open(srs, oracle.fr_from_ints([7]), oracle.fr_from_ints([5])[0])   # constant: empty quotient
    assert oracle.fr_to_ints(v0) == [7] and tsgpu.g1_compress(p0).hex() == "00" * 31 + "40"


def test_open_at_zero_and_one(ctx, tsgpu, oracle, srs_small):
    srs, ref = srs_small
    poly = oracle.chacha_fr_rand(seed_bytes(66), 600)
    for zi in (0, 1):
        z = oracle.fr_from_ints([zi])[0]
        value, proof = tsgpu.KZGCommitment.open(srs, poly, z)
        v_ref, q_ref = oracle.kzg_value_quotient(poly, z)
        assert (value == v_ref).all()
        assert oracle.g1_compress(proof) == oracle.g1_compress(oracle.msm_pippenger(oracle.g1_batch_to_affine(ref[:599]), q_ref))


@pytest.mark.parametrize("n", [1100, 2000, 4096, 4097])
def test_commit_table_mode_matches_oracle_msm(ctx, tsgpu, oracle, srs_small, n):
    """polynomials covering >= 1/4 of the SRS run on the precomputed window tables (one shared bucket set, csrc/msm.cu);
    same group element as the CPU Pippenger over the plain points"""
    srs, ref = srs_small
    poly = oracle.chacha_fr_rand(seed_bytes(n % 251), n).reshape(n, 4)
    got = tsgpu.KZGCommitment.commit(srs, poly)
    want = oracle.msm_pippenger(oracle.g1_batch_to_affine(ref[:n]), poly)
    assert tsgpu.g1_compress(got) == oracle.g1_compress(want)


def test_commit_without_window_tables(ctx, tsgpu, oracle, srs_small):
    """tuning msm_tables = 0: SRS handles built without tables use per-window bucket sets; identical commitments"""
    srs, ref = srs_small
    tau, _ = oracle.setup_scalars()
    try:
        ctx.set_tuning("msm_tables", 0)
        plain = ctx.srs_generate(tau, 3000)
    finally:
        ctx.set_tuning("msm_tables", 1)
    for n in (1, 700, 3000):
        poly = oracle.chacha_fr_rand(seed_bytes(n % 250 + 1), n).reshape(n, 4)
        assert tsgpu.g1_compress(tsgpu.KZGCommitment.commit(plain, poly)) == tsgpu.g1_compress(tsgpu.KZGCommitment.commit(srs, poly))


def test_commit_edge_scalars_table_mode(ctx, tsgpu, oracle, srs_small):
    """heavy buckets in table mode: equal scalars put every point of a window into one bucket (chunked + tree-merged)"""
    srs, ref = srs_small
    n = 4000
    aff = oracle.g1_batch_to_affine(ref[:n])
    for ints in ([1] * n, [123456789] * n, [oracle.R_MOD - 1] * n, [(1 << 253) + 5] * n, [i % 7 for i in range(n)]):
        poly = oracle.fr_from_ints(ints)
        got = tsgpu.KZGCommitment.commit(srs, poly)
        assert tsgpu.g1_compress(got) == oracle.g1_compress(oracle.msm_pippenger(aff, poly))


@pytest.mark.parametrize("chunk", [1, 2, 5, 16, 64])
def test_commit_every_chunk_length_and_merge_path(ctx, tsgpu, oracle, srs_small, chunk, monkeypatch):
    """work-item length (entries one accumulation thread adds; chosen per pass from the bucket count, csrc/msm.cu msm_scratch_bytes) forced through TSGPU_MSM_CHUNK:
    buckets split into up to MSM_SERIAL_MERGE chunks are merged by one thread per bucket (k_msm_merge_serial), up to MSM_BLOCK_MERGE by one block per bucket
    (k_msm_merge_heavy), more by the cooperative tree (k_msm_merge_chunks);
    every combination returns the oracle's group element - uniform scalars, seven distinct scalars (heavy buckets) and short scalars"""
    monkeypatch.setenv("TSGPU_MSM_CHUNK", str(chunk))
    srs, ref = srs_small
    n = 4096
    aff = oracle.g1_batch_to_affine(ref[:n])
    cases = [oracle.chacha_fr_rand(seed_bytes(77), n).reshape(n, 4), oracle.fr_from_ints([i % 7 for i in range(n)]), oracle.fr_from_ints([(i * 2654435761) % (1 << 40) for i in range(n)]),
             oracle.fr_from_ints([123456789] * n)]   # one bucket per window holds every point: 4096 one-entry chunks run the cooperative tree, 2048 and fewer the block-per-bucket merge
    for poly in cases:
        got = tsgpu.KZGCommitment.commit(srs, poly)
        assert tsgpu.g1_compress(got) == oracle.g1_compress(oracle.msm_pippenger(aff, poly))
