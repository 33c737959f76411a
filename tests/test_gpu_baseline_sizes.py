"""GPU vs the CPU oracle at the FULL sizes of BASELINE.json (configs 2-5), bit for bit, on the inputs SURVEY 8(d) defines - the same
generators bench.py times.  The oracle runs its fast tier (NTT interpolation, Pippenger, table-folding sum-check; all host threads) and
its OWN setup_g1_powers: nothing the GPU produced is fed back into the checker except where a test says so.

    C2  Twist::prove, 2^16 cells, 2^20 ops, setup_params(18): distributions A and B, and B with field-sized values   src/twist.rs:107-252
    C3  Shout::prove, 2^20-entry table, 2^22 lookups, setup_params(20)                                               src/shout.rs:97-222
    C4  SumCheck::prove over eq x one-hot, 2^26 entries per table, 26 rounds                                          src/sumcheck.rs:56-110
    C5  KZGCommitment::commit of 2^24 uniform scalars over g1_powers[0..2^24)                                         src/commitments.rs:162-180

Budget (16 host cores): ~25 s per 2^20-op oracle proof, ~2 min for the C3 proof, ~1 min each for C4 / C5 - the module stays well inside
the driver's 1200 s."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import bench  # noqa: E402  (trace generators shared with the benchmark)

pytestmark = pytest.mark.gpu
PKG = "multilinear-map-cryptography_b200"


@pytest.fixture(scope="module")
def oracle_powers(oracle):
    """g1_powers[0 .. 2^22] of setup_params by the ORACLE (src/utils.rs:89-96), Jacobian; its first 2^20 + 1 entries are setup_params(18)'s"""
    return oracle.setup_g1_powers((1 << 22) + 1, fast=True)


@pytest.fixture(scope="module")
def params18(ctx, tsgpu):
    return tsgpu.setup_params(ctx, 18)


def _affine(oracle, jac):
    return oracle.g1_batch_to_affine(np.ascontiguousarray(jac))


def test_c2_srs_of_setup_params_18_equals_the_oracle_setup(ctx, tsgpu, oracle, params18, oracle_powers):
    """all 2^20 + 1 points of the device-generated SRS (k_tau_powers + k_fixed_base_mul + k_batch_to_affine) against the oracle's own setup"""
    pp, _ = params18
    n = (1 << 20) + 1
    assert len(pp.srs) == n
    got = _affine(oracle, pp.srs.download())
    want = _affine(oracle, oracle_powers[:n])
    assert (got == want).all()


@pytest.mark.parametrize("dist", ["A", "B", "B-wide"])
def test_c2_twist_2p20_ops_bytes_equal_the_oracle(ctx, tsgpu, oracle, params18, oracle_powers, dist):
    """the benchmarked proof itself: SURVEY 8(d) distribution A (src/benchmarks.rs:88-99), B (ChaCha20 seed [2; 32]) and B with the written
    values replaced by uniform field elements (no short-scalar commit pass).  Oracle input: its own powers, not the device SRS."""
    pp, vp = params18
    n = 1 << 20
    if dist == "A":
        addr, vals_u64, isw = bench.trace_ref_pattern(20, 16)
        vals = tsgpu.fe_vec(vals_u64)
    else:
        addr, vals_u64, isw = bench.trace_random(20, 16, oracle.chacha_u64(bytes([2]) * 32, 3 * n))
        vals = tsgpu.fe_vec(vals_u64)
        if dist == "B-wide":
            wide = oracle.chacha_fr_rand(bytes([6]) * 32, n).reshape(n, 4)
            src = bench.simulate_memory(addr, isw, np.arange(1, n + 1, dtype=np.uint64))
            vals = np.where((src > 0)[:, None], wide[np.maximum(src, 1).astype(np.int64) - 1], np.uint64(0))
    twist = tsgpu.Twist.new(pp)
    proof = twist.prove_arrays(addr, vals, isw)
    want, _ = oracle.twist_prove(oracle_powers[:n + 1], n, addr, vals, isw, fast=True)
    assert proof.to_bytes() == want
    assert twist.verify(proof, vp)


def test_c3_shout_2p20_table_2p22_lookups_bytes_equal_the_oracle(ctx, tsgpu, oracle, oracle_powers):
    """config 3 at full size, and with it the 2^22 + 1 points of setup_params(20)"""
    T, L = 1 << 20, 1 << 22
    pp, vp = tsgpu.setup_params(ctx, 20)
    assert len(pp.srs) == L + 1
    assert (_affine(oracle, pp.srs.download()) == _affine(oracle, oracle_powers)).all()
    i = np.arange(T, dtype=np.uint64)
    entries = tsgpu.fe_vec(i * i)                                             # src/benchmarks.rs:167-169
    idx = oracle.chacha_u64(bytes([3]) * 32, L) % np.uint64(T)
    shout = tsgpu.Shout.new(pp)
    proof = shout.prove_arrays(entries, idx)
    want, _ = oracle.shout_prove(oracle_powers, L, entries, idx, fast=True)
    assert proof.to_bytes() == want
    assert shout.verify(proof, vp)
    pp.free()


def test_c4_sumcheck_2p26_round_polynomials_equal_the_oracle(ctx, tsgpu, oracle):
    """config 4 at full size: A = eq(w, .), B = one-hot 2^10 x 2^16 (seed [4; 32]); the device builds its tables with its own generators, the
    oracle builds them on the host (eq table by its own routine, one-hot in numpy); all 26 round polynomials, the challenges and the final
    evaluation must agree.  Default (deterministic round 0) and opt-in deferred form."""
    nv, logK = 26, 10
    rows = 1 << (nv - logK)
    w, addr = oracle.chacha_fr_then_u64(bytes([4]) * 32, nv, rows)
    addr = addr % np.uint64(1 << logK)
    hostA = oracle.eq_table(w.reshape(nv, 4))
    hostB = np.zeros((1 << nv, 4), dtype=np.uint64)
    one = oracle.fr_from_ints([1])[0]
    hostB[np.arange(rows, dtype=np.int64) * (1 << logK) + addr.astype(np.int64)] = one
    # claimed sum = sum_j A[j K + addr_j]
    picked = hostA[np.arange(rows, dtype=np.int64) * (1 << logK) + addr.astype(np.int64)]
    claimed = oracle.fr_from_ints([sum(oracle.fr_to_ints(picked)) % oracle.R_MOD])[0]
    ref = oracle.sumcheck_prove_product([hostA, hostB], claimed, mode="tables")
    del hostA, hostB
    A = ctx.table_eq(w.reshape(nv, 4)); B = ctx.table_one_hot_rows(addr, logK, nv)
    try:
        for flag in (0, 1):
            ctx.set_tuning("deferred_claim_check", flag)
            proof, chals, finals = tsgpu.SumCheck(nv, claimed).prove_product(ctx, [A.clone(), B.clone()], tsgpu.Transcript(), return_aux=True)
            assert (proof.round_polynomials == ref["round_polynomials"]).all()
            assert (proof.final_evaluation == ref["final_evaluation"]).all()
            assert (chals == ref["challenges"]).all() and (finals == ref["finals"]).all()
    finally:
        ctx.set_tuning("deferred_claim_check", 0)


def test_c5_msm_2p24_commitment_equals_the_oracle_pippenger(ctx, tsgpu, oracle, oracle_powers):
    """config 5 at full size: 2^24 uniform full-width scalars (Fr::rand, seed [5; 32]) over g1_powers[0 .. 2^24).  The device SRS is checked
    against the oracle first - its first 2^22 + 1 points against the oracle's own setup, the rest at 64 sampled exponents against
    tau^i G by the oracle's double-and-add - and then serves as the oracle's base points."""
    n = 1 << 24
    tau, _ = oracle.setup_scalars()
    srs = ctx.srs_generate(tau, n)
    pts = _affine(oracle, srs.download())
    m = oracle_powers.shape[0]
    assert (pts[:m] == _affine(oracle, oracle_powers)).all()
    t = oracle.fr_to_ints(tau.reshape(1, 4))[0]
    G = oracle.g1_generator()
    sample = [m, n - 1, n - 2, 1 << 23, (1 << 23) + 1] + [int(x) for x in np.random.default_rng(24).integers(m, n, size=59)]
    for i in sample:
        want = oracle.g1_mul(G, oracle.fr_from_ints([pow(t, i, oracle.R_MOD)])[0])
        assert (_affine(oracle, want.reshape(1, 12))[0] == pts[i]).all(), i
    scalars = oracle.chacha_fr_rand(bytes([5]) * 32, n).reshape(n, 4)
    got = tsgpu.KZGCommitment.commit(srs, ctx.poly_upload(scalars))
    want = oracle.msm_pippenger(pts, scalars)
    assert tsgpu.g1_compress(got) == oracle.g1_compress(want)
    srs.free()
