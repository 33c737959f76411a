"""CPU: pins the C++ oracle (oracle/oracle.cpp) against (a) the golden vectors of tests/golden/appendix_c.json
(SURVEY.md Appendix C + full proof bytes from the independent pure-Python restatement), (b) published
known-answer vectors of the primitives, (c) the reference tests' own numeric anchors, and (d) itself:
verbatim tier == fast tier."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import seed_bytes

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "appendix_c.json")))


def _ops_arrays(oracle, ops):
    addr = np.array([o[1] for o in ops], dtype=np.uint64)
    vals = oracle.fr_from_ints([o[2] for o in ops]) if ops else np.empty((0, 4), dtype=np.uint64)
    isw = np.array([o[0] == "W" for o in ops], dtype=np.uint8)
    return addr, vals, isw


def test_primitive_known_answers(oracle):
    import pyref
    assert GOLD["siphash13_empty"] == "0xd1fba762150c532c" == hex(oracle.siphash13(b""))        # Rust DefaultHasher::new().finish()
    assert GOLD["chacha20_zero_key_words"] == ["0xade0b876", "0x903df1a0"]                      # RFC 7539 zero-key keystream
    assert int(oracle.chacha_u64(bytes(32), 1)[0]) == 0x903df1a0ade0b876
    for n in (0, 1, 7, 8, 9, 63, 64, 65, 1000):
        d = bytes((i * 37 + n) & 0xFF for i in range(n))
        assert oracle.siphash13(d) == pyref.siphash13(d)
    rng = pyref.ChaCha20Rng(seed_bytes(9))
    assert oracle.fr_to_ints(oracle.chacha_fr_rand(seed_bytes(9), 80)) == [pyref.fr_rand(rng) for _ in range(80)]
    rng = pyref.ChaCha20Rng(seed_bytes(5))
    assert list(map(int, oracle.chacha_u64(seed_bytes(5), 150))) == [rng.next_u64() for _ in range(150)]


def test_setup_scalars_and_powers(oracle):
    tau, seed = oracle.setup_scalars()
    assert str(oracle.fr_to_ints(tau)[0]) == GOLD["tau"] and seed.hex() == GOLD["fiat_shamir_seed"]
    assert hex(oracle.limbs_to_ints(tau)[0]) == GOLD["tau_montgomery_limbs_hex"]
    pw = oracle.setup_g1_powers(33, fast=False)
    assert oracle.g1_compress(pw[1]).hex() == GOLD["g1_powers_1"] and oracle.g1_compress(pw[32]).hex() == GOLD["g1_powers_32"]
    assert oracle.g1_compress(oracle.setup_g1_powers(33, fast=True)) == oracle.g1_compress(pw)
    assert oracle.g1_compress(pw[0]).hex() == "01" + "00" * 31          # generator (1, 2)


def test_transcript_golden(oracle):
    t = oracle.Transcript()
    t.append_field_element(b"test", oracle.fr_from_ints([123])[0])
    assert str(oracle.fr_to_ints(t.challenge_field_element(b"challenge"))[0]) == GOLD["transcript_test_challenge"]


@pytest.mark.parametrize("case", ["twist_demo", "twist_readme", "twist_empty"])
@pytest.mark.parametrize("fast", [False, True])
def test_twist_proof_bytes_golden(oracle, case, fast):
    g = GOLD[case]
    pw = oracle.setup_g1_powers(40, fast=True)
    addr, vals, isw = _ops_arrays(oracle, g["ops"])
    by, z = oracle.twist_prove(pw, 4 << g["log_size"], addr, vals, isw, fast=fast)
    assert by.hex() == g["proof_hex"]
    if "z" in g:
        assert str(oracle.fr_to_ints(z)[0]) == g["z"]


@pytest.mark.parametrize("case", ["shout_demo", "shout_readme", "shout_no_lookups"])
@pytest.mark.parametrize("fast", [False, True])
def test_shout_proof_bytes_golden(oracle, case, fast):
    g = GOLD[case]
    pw = oracle.setup_g1_powers(40, fast=True)
    by, _ = oracle.shout_prove(pw, 4 << g["log_size"], oracle.fr_from_ints(g["entries"]), np.array(g["lookups"], dtype=np.uint64), fast=fast)
    assert by.hex() == g["proof_hex"]


@pytest.mark.parametrize("mode", ["closure", "tables"])
def test_sumcheck_c3_golden(oracle, mode):
    g = GOLD["sumcheck_c3"]
    r = oracle.sumcheck_prove_product([oracle.fr_from_ints(g["A"]), oracle.fr_from_ints(g["B"])], oracle.fr_from_ints([g["claimed_sum"]])[0], mode=mode)
    assert [[str(c) for c in oracle.fr_to_ints(rp)] for rp in r["round_polynomials"]] == g["round_polynomials"]
    assert [str(c) for c in oracle.fr_to_ints(r["challenges"])] == g["challenges"]
    assert str(oracle.fr_to_ints(r["final_evaluation"])[0]) == g["final_evaluation"]


def test_reference_numeric_anchors(oracle):
    """src/commitments.rs:509 f(5)=86; tests/polynomial_tests.rs:93-131,191-208; src/polynomials.rs:441-442 lt rows"""
    v, q = oracle.kzg_value_quotient(oracle.fr_from_ints([1, 2, 3]), oracle.fr_from_ints([5])[0])
    assert oracle.fr_to_ints(v) == [86] and oracle.fr_to_ints(q) == [17, 3]
    xs = oracle.fr_from_ints([0, 1, 2]); ys = oracle.fr_from_ints([0, 1, 4])
    assert oracle.fr_to_ints(oracle.lagrange_interpolate(xs, ys)) == [0, 0, 1]
    half = pow(2, -1, oracle.R_MOD)
    e = oracle.mle_evaluate(oracle.fr_from_ints([1, 2, 3, 4]), oracle.fr_from_ints([half, half]))
    assert oracle.fr_to_ints(e)[0] == 10 * pow(4, -1, oracle.R_MOD) % oracle.R_MOD
    assert oracle.fr_to_ints(oracle.mle_partial_evaluate(oracle.fr_from_ints([1, 2, 3, 4]), oracle.fr_from_ints([1]))) == [2, 4]
    lt = oracle.fr_to_ints(oracle.lt_table(2))
    # index = a | b << 2; lt decides on the first differing LOW bit: lt(a=1 (bits 1,0), b=2 (bits 0,1)) = 0, lt(2,1) = 1
    assert lt[1 | (2 << 2)] == 0 and lt[2 | (1 << 2)] == 1 and lt[0 | (3 << 2)] == 1 and lt[3 | (3 << 2)] == 0


def test_fast_tier_equals_verbatim_tier(oracle):
    for n in (1, 2, 8, 64):
        vals = oracle.chacha_fr_rand(seed_bytes(n), n).reshape(n, 4)
        xs = oracle.fr_from_ints(list(range(n)))
        assert (oracle.interpolate_iota_fast(vals) == oracle.lagrange_interpolate(xs, vals)).all()
    pw = oracle.setup_g1_powers(300, fast=True)
    poly = oracle.chacha_fr_rand(seed_bytes(77), 300)
    assert oracle.g1_equal(oracle.kzg_commit(pw, poly), oracle.msm_pippenger(oracle.g1_batch_to_affine(pw), poly))
    ev = oracle.chacha_fr_rand(seed_bytes(3), 256); pt = oracle.chacha_fr_rand(seed_bytes(4), 8)
    assert (oracle.mle_evaluate(ev, pt) == oracle.mle_evaluate(ev, pt, fold=True)).all()
    fx = oracle.chacha_fr_rand(seed_bytes(6), 3)
    assert (oracle.mle_partial_evaluate(ev, fx) == oracle.mle_partial_evaluate(ev, fx, fold=True)).all()
    # random 100-op trace: verbatim prover (O(n^3) interpolation, serial commit, closure sum-check) == fast prover
    rng = np.random.default_rng(1)
    addr = rng.integers(0, 16, size=100).astype(np.uint64)
    vals = oracle.chacha_fr_rand(seed_bytes(8), 100); isw = rng.integers(0, 2, size=100).astype(np.uint8)
    b1, _ = oracle.twist_prove(pw, 512, addr, vals, isw, fast=False)
    b2, _ = oracle.twist_prove(pw, 512, addr, vals, isw, fast=True)
    assert b1 == b2


def test_constraint_sumchecks_are_consistent_on_the_reference_sumcheck(oracle):
    """The identities behind the non-parity constraint mode, checked by the reference's own SumCheck::prove (closure form: every round's
    g(0) + g(1) check, src/sumcheck.rs:77-84): rv~(r) = sum_x ra~(x, r) Val~(x) for lookups, and for memory
    sum_j eq(r, j)[read_j] v_j = sum_{x,j} eq(r, j)[read_j] ra(x, j) Val(x, j) together with Val~(x*, j*) = sum_j' Inc_j' eq(x*, a_j') LT~(j', j*).
    A wrong returned value must break round 0."""
    rng = np.random.default_rng(11)
    entries = oracle.chacha_fr_rand(bytes([3]) * 32, 8).reshape(8, 4)
    idx = rng.integers(0, 8, size=5).astype(np.uint64)
    vals = entries[idx.astype(np.int64)]
    claim, ref = oracle.shout_read_check_prove(entries, idx, vals, mode="closure")
    assert ref["round_polynomials"].shape == (3, 4, 4) and ref["round_polynomials"].any()
    bad = vals.copy(); bad[2] = entries[(int(idx[2]) + 1) % 8]
    with pytest.raises(oracle.SumCheckError):
        oracle.shout_read_check_prove(entries, idx, bad, mode="closure")
    # memory: 4 cells, W(0,a) W(1,b) R(0) W(0,c) R(0) R(1) R(3)
    v = oracle.fr_to_ints(oracle.chacha_fr_rand(bytes([4]) * 32, 3))
    addr = np.array([0, 1, 0, 0, 0, 1, 3], dtype=np.uint64)
    isw = np.array([1, 1, 0, 1, 0, 0, 0], dtype=np.uint8)
    mvals = oracle.fr_from_ints([v[0], v[1], v[0], v[2], v[2], v[1], 0])
    c1, c2, r1, r2 = oracle.twist_memory_check_prove(addr, mvals, isw, 4, mode="closure")
    assert r1["round_polynomials"].shape == (5, 4, 4) and r2["round_polynomials"].shape == (3, 4, 4)
    wrong = oracle.fr_from_ints([v[0], v[1], v[0], v[2], v[0], v[1], 0])          # the second read of cell 0 returns the overwritten value
    with pytest.raises(oracle.SumCheckError):
        oracle.twist_memory_check_prove(addr, wrong, isw, 4, mode="closure")
    # LT~ at a boolean point is the indicator of the natural order
    assert oracle.lt_point_ints([1, 0, 1], 3) == [1 if a < 5 else 0 for a in range(8)]
