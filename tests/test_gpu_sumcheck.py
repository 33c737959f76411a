"""GPU parity: sum-check round kernels (through the C ABI) vs the CPU oracle, bit-exact.

Reference: SumCheck::prove / compute_round_polynomial (src/sumcheck.rs:56-110,156-207) called with
f(v) = prod_t mle_t.evaluate(v)."""
import numpy as np
import pytest

from conftest import seed_bytes

pytestmark = pytest.mark.gpu


def drive_rounds(ctx, oracle, host_tables, fused=True):
    """Run the device sum-check with the oracle's transcript supplying challenges.
    Returns (round evals list, challenges, finals)."""
    tabs = [ctx.table_upload(t) for t in host_tables]
    nv = tabs[0].num_vars
    sc = ctx.sumcheck(tabs)
    tr = oracle.Transcript()
    xs = oracle.fr_from_ints([0, 1, 2, 3])
    evals_all, chals, coeffs_all = [], [], []
    ev = sc.round_eval() if nv else None
    for rnd in range(nv):
        evals_all.append(ev.copy())
        coeffs = oracle.lagrange_interpolate(xs, ev)
        coeffs_all.append(coeffs)
        tr.append_field_elements(f"sumcheck_round_{rnd}".encode(), coeffs)
        r = tr.challenge_field_element(f"sumcheck_challenge_{rnd}".encode())
        chals.append(r)
        if rnd + 1 < nv:
            if fused:
                ev = sc.bind_eval(r)
            else:
                sc.bind(r)
                ev = sc.round_eval()
        else:
            sc.bind(r)
    finals = sc.final()
    sc.end()
    return np.array(coeffs_all).reshape(nv, 4, 4), np.array(chals).reshape(nv, 4), finals


@pytest.mark.parametrize("d", [1, 2, 3])
@pytest.mark.parametrize("nv", [1, 2, 3, 5, 8, 11])
@pytest.mark.parametrize("fused", [True, False])
def test_sumcheck_product_matches_oracle(ctx, oracle, d, nv, fused):
    n = 1 << nv
    tables = [oracle.chacha_fr_rand(seed_bytes(10 * d + t + nv), n) for t in range(d)]
    ref_closure_ok = nv <= 5
    ints = [oracle.fr_to_ints(t) for t in tables]
    claimed = 0
    for i in range(n):
        p = 1
        for t in range(d):
            p = p * ints[t][i] % oracle.R_MOD
        claimed = (claimed + p) % oracle.R_MOD
    claimed_m = oracle.fr_from_ints([claimed])[0]
    ref = oracle.sumcheck_prove_product(tables, claimed_m, mode="closure" if ref_closure_ok else "tables")
    coeffs, chals, finals = drive_rounds(ctx, oracle, tables, fused=fused)
    assert (coeffs == ref["round_polynomials"]).all()
    assert (chals == ref["challenges"]).all()
    fe = oracle.fr_from_ints([1])[0]
    for t in range(d):
        fe = oracle.field_binop("fr", "mul", fe, finals[t])[0]
    assert (fe == ref["final_evaluation"]).all()
    tab_ref = oracle.sumcheck_prove_product(tables, claimed_m, mode="tables")
    assert (finals == tab_ref["finals"]).all()


def test_appendix_c3_vector(ctx, oracle):
    """A=[1..8], B=[3,1,4,1,5,9,2,6], claimed sum 162."""
    A = oracle.fr_from_ints(list(range(1, 9)))
    B = oracle.fr_from_ints([3, 1, 4, 1, 5, 9, 2, 6])
    coeffs, chals, finals = drive_rounds(ctx, oracle, [A, B])
    assert oracle.fr_to_ints(coeffs[0]) == [54, 51, 3, 0]
    assert oracle.fr_to_ints(chals)[0] == 21125437990100363807064869691380599404649938677815818472735327310510785473320
    fe = oracle.fr_to_ints(oracle.field_binop("fr", "mul", finals[0], finals[1]))[0]
    assert fe == 359745214182377975469500176028792295272184875172441352787133304826151130394


def test_large_round_linearity(ctx, oracle):
    """2^20 entries, d=2: g(0)+g(1) of round k must equal g_{k-1}(r_{k-1}) - the size-independent
    sum-check invariant - and the final product must equal the last round polynomial at the last challenge."""
    nv = 20
    n = 1 << nv
    A = oracle.chacha_fr_rand(seed_bytes(91), n)
    B = oracle.chacha_fr_rand(seed_bytes(92), n)
    coeffs, chals, finals = drive_rounds(ctx, oracle, [A, B])
    zero, one = oracle.fr_from_ints([0, 1])
    claimed = oracle.field_binop("fr", "add", oracle.horner(coeffs[0], zero), oracle.horner(coeffs[0], one))[0]
    ref = oracle.sumcheck_prove_product([A, B], claimed, mode="tables")
    assert (coeffs == ref["round_polynomials"]).all()
    assert (finals == ref["finals"]).all()


@pytest.mark.parametrize("nv", [12, 17, 18])
def test_warp_prefetch_path_matches_oracle(ctx, tsgpu, oracle, nv):
    """d = 2 with the warp-private prefetch kernels (cp.async.bulk into one shared-memory slot per warp) switched on for every launch
    of at least 2^10 positions, the later rounds on the plain kernels: the full proof through the C++ host loop (round 0 evaluation
    kernel, then the fused claim form) must equal the oracle's, and so must the round-by-round drive without the claim."""
    tables = [oracle.chacha_fr_rand(seed_bytes(80 + t + nv), 1 << nv) for t in range(2)]
    ctx.set_tuning("prefetch_min_log2", 10)
    try:
        coeffs, chals, finals = drive_rounds(ctx, oracle, tables, fused=False)        # k_round_eval2_pf every round
        zero, one = oracle.fr_from_ints([0, 1])
        claimed = oracle.field_binop("fr", "add", oracle.horner(coeffs[0], zero), oracle.horner(coeffs[0], one))[0]
        proof, ch2, fin2 = tsgpu.SumCheck(nv, claimed).prove_product(ctx, [ctx.table_upload(t) for t in tables], tsgpu.Transcript(), return_aux=True)
    finally:
        ctx.set_tuning("prefetch_min_log2", -1)
    ref = oracle.sumcheck_prove_product(tables, claimed, mode="tables")
    assert (coeffs == ref["round_polynomials"]).all() and (chals == ref["challenges"]).all() and (finals == ref["finals"]).all()
    assert (proof.round_polynomials == ref["round_polynomials"]).all() and (proof.final_evaluation == ref["final_evaluation"]).all()
    assert (ch2 == ref["challenges"]).all() and (fin2 == ref["finals"]).all()


@pytest.mark.parametrize("d,nv", [(1, 4), (2, 6), (3, 5), (2, 12)])
def test_prove_product_host_loop_matches_reference_prove(ctx, tsgpu, oracle, d, nv):
    """The C++ host loop (host/sumcheck_host.cpp) + device rounds == SumCheck::prove with the product closure:
    same round polynomials, same final evaluation, same transcript state afterwards; verify() accepts."""
    n = 1 << nv
    tables = [oracle.chacha_fr_rand(seed_bytes(60 + 5 * d + t), n) for t in range(d)]
    ints = [oracle.fr_to_ints(t) for t in tables]
    claimed = 0
    for i in range(n):
        p = 1
        for t in range(d):
            p = p * ints[t][i] % oracle.R_MOD
        claimed = (claimed + p) % oracle.R_MOD
    claimed_m = oracle.fr_from_ints([claimed])[0]
    otr = oracle.Transcript()
    ref = oracle.sumcheck_prove_product(tables, claimed_m, transcript=otr, mode="closure" if nv <= 6 else "tables")
    tr = tsgpu.Transcript()
    sc = tsgpu.SumCheck(nv, claimed_m)
    proof, chals, finals = sc.prove_product(ctx, [ctx.table_upload(t) for t in tables], tr, return_aux=True)
    assert (proof.round_polynomials == ref["round_polynomials"]).all()
    assert (proof.final_evaluation == ref["final_evaluation"]).all()
    assert (chals == ref["challenges"]).all()
    # both transcripts must now produce the same next challenge
    assert (tr.challenge_field_element(b"next") == otr.challenge_field_element(b"next")).all()
    ok, ch2 = sc.verify(proof, tsgpu.Transcript())
    assert ok and (ch2 == chals).all()


def test_wrong_claimed_sum_is_rejected_in_round_zero(ctx, tsgpu, oracle):
    """sumcheck.rs:77-84: Err(SumCheck("Round 0 consistency check failed"))"""
    tables = [oracle.chacha_fr_rand(seed_bytes(70 + t), 64) for t in range(2)]
    with pytest.raises(tsgpu.TwistAndShoutError) as e:
        tsgpu.SumCheck(6, oracle.fr_from_ints([12345])[0]).prove_product(ctx, [ctx.table_upload(t) for t in tables], tsgpu.Transcript())
    assert e.value.variant == "SumCheck" and "Round 0 consistency check failed" in str(e.value)
    with pytest.raises(oracle.SumCheckError):
        oracle.sumcheck_prove_product(tables, oracle.fr_from_ints([12345])[0], mode="closure")


@pytest.mark.parametrize("nv", [1, 2, 7, 13])
def test_round_zero_check_default_and_deferred_forms_behave_like_the_reference(ctx, tsgpu, oracle, nv):
    """d = 2.  Default: the reference's deterministic round-0 check (sumcheck.rs:77-84) - a wrong claim fails before anything is appended and the
    tables are untouched.  Opt-in "deferred_claim_check": round 0 in the claim form, the claimed sum checked at the end.  Both must give
    Err(SumCheck("Round 0 consistency check failed")) for a wrong claim, leave the transcript exactly as the reference does (unchanged), and
    give the oracle's proof for the right claim."""
    n = 1 << nv
    tables = [oracle.chacha_fr_rand(seed_bytes(90 + t + nv), n) for t in range(2)]
    ai, bi = oracle.fr_to_ints(tables[0]), oracle.fr_to_ints(tables[1])
    claimed = sum(x * y for x, y in zip(ai, bi)) % oracle.R_MOD
    good = oracle.fr_from_ints([claimed])[0]
    ref = oracle.sumcheck_prove_product(tables, good, mode="tables")
    for flag in (0, 1):
        ctx.set_tuning("deferred_claim_check", flag)
        try:
            for delta in (1, oracle.R_MOD - 1, 123456789):
                bad = oracle.fr_from_ints([(claimed + delta) % oracle.R_MOD])[0]
                tr = tsgpu.Transcript()
                tr.append_field_element(b"before", tsgpu.fe(7))
                mark = tr.state_len
                dev = [ctx.table_upload(t) for t in tables]
                with pytest.raises(tsgpu.TwistAndShoutError) as e:
                    tsgpu.SumCheck(nv, bad).prove_product(ctx, dev, tr)
                assert e.value.variant == "SumCheck" and e.value.message == "Round 0 consistency check failed"
                assert tr.state_len == mark                               # the next challenge is what the reference would draw
                otr = oracle.Transcript(); otr.append_field_element(b"before", tsgpu.fe(7))
                assert (tr.challenge_field_element(b"next") == otr.challenge_field_element(b"next")).all()
                if flag == 0:                                             # deterministic form: the caller's tables are still whole
                    assert dev[0].num_vars == nv and (dev[0].download() == tables[0].reshape(-1, 4)).all() and (dev[1].download() == tables[1].reshape(-1, 4)).all()
            proof = tsgpu.SumCheck(nv, good).prove_product(ctx, [ctx.table_upload(t) for t in tables], tsgpu.Transcript())
        finally:
            ctx.set_tuning("deferred_claim_check", 0)
        assert (proof.round_polynomials == ref["round_polynomials"]).all() and (proof.final_evaluation == ref["final_evaluation"]).all()


def test_a_thousand_random_wrong_claims_fail_identically_in_both_forms(ctx, tsgpu, oracle):
    """10^3 random wrong claims on a 2^4-entry product: the deterministic and the deferred form return the same error and leave the same transcript state"""
    nv = 4
    tables = [oracle.chacha_fr_rand(seed_bytes(200 + t), 1 << nv) for t in range(2)]
    claims = oracle.chacha_fr_rand(seed_bytes(203), 1000).reshape(-1, 4)
    want = tsgpu.Transcript(); want.append_field_element(b"before", tsgpu.fe(1)); want_next = want.challenge_field_element(b"next")
    try:
        for flag in (0, 1):
            ctx.set_tuning("deferred_claim_check", flag)
            for c in claims:
                tr = tsgpu.Transcript(); tr.append_field_element(b"before", tsgpu.fe(1))
                with pytest.raises(tsgpu.TwistAndShoutError) as e:
                    tsgpu.SumCheck(nv, c).prove_product(ctx, [ctx.table_upload(t) for t in tables], tr)
                assert e.value.message == "Round 0 consistency check failed"
                assert (tr.challenge_field_element(b"next") == want_next).all()
    finally:
        ctx.set_tuning("deferred_claim_check", 0)


def test_reference_integration_x1_times_x2(ctx, tsgpu, oracle):
    """tests/integration_tests.rs:263-288 and src/sumcheck.rs:221-245: f(x1,x2) = x1*x2 sums to 1 over {0,1}^2.
    As tables: A[i] = bit0(i), B[i] = bit1(i)."""
    A = oracle.fr_from_ints([0, 1, 0, 1]); B = oracle.fr_from_ints([0, 0, 1, 1])
    one = oracle.fr_from_ints([1])[0]
    sc = tsgpu.SumCheck(2, one)
    proof = sc.prove_product(ctx, [ctx.table_upload(A), ctx.table_upload(B)], tsgpu.Transcript(bytes([42]) * 32))
    ok, _ = sc.verify(proof, tsgpu.Transcript(bytes([42]) * 32))
    assert ok


@pytest.mark.parametrize("nv", [2, 3, 7, 12, 17])
def test_bind_eval_with_claim_returns_the_same_four_values(ctx, oracle, nv):
    """tsgpu_sc_bind_eval_claim derives g(1) = claim - g(0): with the true claim g_k(r) the four evaluations equal the fully summed ones"""
    A = oracle.chacha_fr_rand(seed_bytes(nv + 40), 1 << nv).reshape(-1, 4)
    B = oracle.chacha_fr_rand(seed_bytes(nv + 41), 1 << nv).reshape(-1, 4)
    rs = oracle.chacha_fr_rand(seed_bytes(nv + 42), nv).reshape(-1, 4)
    full = ctx.sumcheck([ctx.table_upload(A), ctx.table_upload(B)])
    fast = ctx.sumcheck([ctx.table_upload(A), ctx.table_upload(B)])
    ev = full.round_eval()
    assert (fast.round_eval() == ev).all()
    for k in range(nv - 1):
        # claim of the next round = g_k(r_k): Lagrange evaluation of the cubic through ev at r_k, done by the oracle's field ops
        c = ints = oracle.fr_to_ints(ev)
        r = oracle.fr_to_ints(rs[k:k + 1])[0]
        P = oracle.R_MOD
        g = 0
        for i in range(4):
            num, den = 1, 1
            for j in range(4):
                if j != i:
                    num = num * (r - j) % P; den = den * (i - j) % P
            g = (g + ints[i] * num * pow(den, -1, P)) % P
        claim = oracle.fr_from_ints([g])
        ev = full.bind_eval(rs[k:k + 1])
        got = fast.bind_eval(rs[k:k + 1], claim=claim)
        assert (got == ev).all(), (nv, k)


def _claim_of(oracle, coeffs, r):
    """g(r) by Horner on the oracle's field: the claim of the next round"""
    return oracle.horner(coeffs, r)


@pytest.mark.parametrize("nv", [2, 3, 7, 11, 12, 14])
def test_persistent_tail_rounds_match_oracle_and_the_per_round_kernels(ctx, tsgpu, oracle, nv):
    """d = 2: once the tables fit one CTA's shared memory (<= 2^11 entries each after the fold) ONE resident kernel runs the remaining rounds and
    talks to the host through a mapped mailbox (tuning "sc_tail", default on).  Same proof with the tail on and off, equal to the oracle's;
    and a caller that mixes the claim form with plain bind / round_eval / bind_eval calls (which make the kernel hand the tables back)
    gets the same values as the per-round kernels."""
    tables = [oracle.chacha_fr_rand(seed_bytes(120 + t + nv), 1 << nv) for t in range(2)]
    ai, bi = oracle.fr_to_ints(tables[0]), oracle.fr_to_ints(tables[1])
    claimed = oracle.fr_from_ints([sum(x * y for x, y in zip(ai, bi)) % oracle.R_MOD])[0]
    ref = oracle.sumcheck_prove_product(tables, claimed, mode="tables")
    launches = {}
    try:
        for flag in (1, 0):
            ctx.set_tuning("sc_tail", flag)
            l0 = ctx.launch_count
            proof, chals, finals = tsgpu.SumCheck(nv, claimed).prove_product(ctx, [ctx.table_upload(t) for t in tables], tsgpu.Transcript(), return_aux=True)
            launches[flag] = ctx.launch_count - l0
            assert (proof.round_polynomials == ref["round_polynomials"]).all() and (proof.final_evaluation == ref["final_evaluation"]).all()
            assert (chals == ref["challenges"]).all() and (finals == ref["finals"]).all()
        if nv >= 3:
            assert launches[1] < launches[0]                     # the tail replaces one launch per round by one launch in all
        # round-stepped drive mixing the entry points: claim form for two rounds, then a plain bind + round_eval, then bind_eval without a claim
        ctx.set_tuning("sc_tail", 1)
        sc = ctx.sumcheck([ctx.table_upload(t) for t in tables])
        sc.exclusive(True)
        xs = oracle.fr_from_ints([0, 1, 2, 3])
        ev = sc.round_eval()
        for rnd in range(nv):
            coeffs = oracle.lagrange_interpolate(xs, ev)
            assert (coeffs == ref["round_polynomials"][rnd]).all(), rnd
            r = ref["challenges"][rnd]
            if rnd + 1 == nv:
                sc.bind(r)
            elif rnd % 3 == 2:
                sc.bind(r); ev = sc.round_eval()
            elif rnd % 3 == 1 and rnd > 2:
                ev = sc.bind_eval(r)
            else:
                ev = sc.bind_eval(r, claim=_claim_of(oracle, coeffs, r))
        assert (sc.final() == ref["finals"]).all()
        sc.end()
        # a handle abandoned while the kernel is resident (error path of a caller): end() must release the GPU
        if nv >= 3:
            sc = ctx.sumcheck([ctx.table_upload(t) for t in tables])
            sc.exclusive(True)
            ev = sc.round_eval()
            coeffs = oracle.lagrange_interpolate(xs, ev)
            sc.bind_eval(ref["challenges"][0], claim=_claim_of(oracle, coeffs, ref["challenges"][0]))
            sc.end()
            assert (ctx.mle_evaluate(tables[0], ref["challenges"]) == oracle.mle_evaluate(tables[0], ref["challenges"], fold=True)).all()   # the stream is free again
    finally:
        ctx.set_tuning("sc_tail", 1)
