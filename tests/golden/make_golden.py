"""Regenerates tests/golden/appendix_c.json from the independent pure-Python restatement (oracle/pyref.py).

The reference ships no golden vectors and cannot be built here (Rust, no toolchain), so the
pinned values are: (1) the SURVEY.md Appendix C cross-check vectors, transcribed below as SURVEY_* constants and
asserted against what pyref computes, and (2) full canonical proof bytes for the reference's own demo inputs
(README quick start and examples/demo.rs), produced by pyref's loop-for-loop O(n^3) restatement.
Run:  python tests/golden/make_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import pyref as P  # noqa: E402

SURVEY_TAU = 8122400061003384056342786174786292760507720762713395576999582764556933302441
SURVEY_SEED = "7bfcfd7544b1078dda397cef45df2e6de498746805081ebc8fb90ad04eba9d02"
SURVEY_G1_1 = "f82610fe9c43824626b034bd432a3a7335eea949272763d214789731a135deaa"
SURVEY_G1_32 = "ace48dfeab869c1bbb618f4ea03e8910e1f917df382545a6cfc012da97bcbaac"
SURVEY_TRANSCRIPT = 13648926573440158680322210633940909009220968087751212041477676025471912345605
SURVEY_TWIST_DEMO = dict(
    address_commitment="bd975589d6f2ede8691ef7e9ca0db40b9e189337547e3fd94404fc49a0731612",
    value_commitment="ff375480179037a41ce77a54486f3a858bc71ddb32c341198c9434ede43360ad",
    z=3376650823654878944670756306812172698111738876918527020609947282539906298609,
    final_evaluations=[15980070657345115799942694709683672561018519327923965335736198485833074678685,
                       7406763215123453977357787948121022732276212498932776477372630137447767779024],
    opening_proofs=["a9616d06cd0ed0537ae110f4eccb91e359d05dfa3cab004af0449f3872cf1816",
                    "0aff02e8661f2c5c968d41e64cc7c8665cb6b6cf833dcc04bbfa048388c94596"])
SURVEY_SHOUT_DEMO = dict(
    table_commitment="7b888bf99c108e9c2fe2335e2f481c85bdedaa21e8b50c7fb8e31e5792ba3202",
    index_commitment="ac7b7853de14af0d9eda46cbd3a8f93b21c8560871a256931c3d74a96241718e",
    z=20777763851659902838829843219060961285863197251027616816986442371813014367890,
    opening_proofs=["5bed70a6518bdd8ff735c364d45689ec556d403c9231217962b1ae6a98d0198f",
                    "9d474549e7da0d742c8203f99e83f72adf95f14a7442e0a4989af8bc1d80d8ae"])
SURVEY_README_TWIST = dict(
    address_commitment="cdfecfeb80caeef10955a942c74bcd530dbe397b3ca0118cbe7beb1b99afb490",
    value_commitment="7485b7fbe07ceee04998d88d4683b99cd0e20864913f992d64072e96130f0220",
    z=6128647445570684818819108835479309113388490358966680049700171301306498851209,
    opening_proofs=["4deea6b5b1dff703c85b63769ec67e9676132a777914003f24757f05f8d35e24",
                    "34d084c1a568f1ed53db943df567761f878ce3097c25ef6e10ebf7a03fb54f2b"])
SURVEY_README_SHOUT = dict(table_commitment="85cd5e645bcce6dec8c497c97bc44d4dc41d399afdbab41b6e8a20089b2ceb04",
                           index_commitment="01" + "00" * 31)
SURVEY_C3 = dict(round0=[54, 51, 3, 0],
                 challenges=[21125437990100363807064869691380599404649938677815818472735327310510785473320,
                             6332574201562144554565502755015533931823701737341446997886563742120444013847,
                             5414186714876751130115553868568506973668294371468978031254075697539025637687],
                 final_evaluation=359745214182377975469500176028792295272184875172441352787133304826151130394)


def main():
    out = {}
    tau, raw, seed = P.setup_tau_and_seed()
    assert tau == SURVEY_TAU and seed.hex() == SURVEY_SEED
    out["tau"] = str(tau); out["tau_montgomery_limbs_hex"] = hex(raw); out["fiat_shamir_seed"] = seed.hex()
    pp3 = P.setup_params(3)
    assert P.g1_compressed(pp3.g1_powers[1]).hex() == SURVEY_G1_1 and P.g1_compressed(pp3.g1_powers[32]).hex() == SURVEY_G1_32
    out["g1_powers_1"] = SURVEY_G1_1; out["g1_powers_32"] = SURVEY_G1_32
    t = P.Transcript(); t.append_field_element(b"test", 123)
    c = t.challenge_field_element(b"challenge")
    assert c == SURVEY_TRANSCRIPT
    out["transcript_test_challenge"] = str(c)
    out["siphash13_empty"] = hex(P.siphash13(b""))
    out["chacha20_zero_key_words"] = [hex(w) for w in P._chacha_block([0] * 8, 0)[:2]]

    # examples/demo.rs: setup_params(3)
    ops = [("W", 0, 42), ("W", 1, 100), ("R", 0, 42), ("R", 1, 100), ("W", 0, 43), ("R", 0, 43)]
    pr = P.twist_prove(pp3, ops)
    assert P.g1_compressed(pr.commitments[0]).hex() == SURVEY_TWIST_DEMO["address_commitment"]
    assert P.g1_compressed(pr.commitments[1]).hex() == SURVEY_TWIST_DEMO["value_commitment"]
    assert pr.z == SURVEY_TWIST_DEMO["z"] and pr.final_evaluations == SURVEY_TWIST_DEMO["final_evaluations"]
    assert [P.g1_compressed(x).hex() for x in pr.opening_proofs] == SURVEY_TWIST_DEMO["opening_proofs"]
    out["twist_demo"] = dict(log_size=3, ops=[[k, a, v] for k, a, v in ops], proof_hex=pr.to_bytes().hex(), z=str(pr.z),
                             address_poly=[str(x) for x in pr.polys[0]])
    sp = P.shout_prove(pp3, [i * i for i in range(8)], [3, 5, 0, 7])
    assert P.g1_compressed(sp.commitments[0]).hex() == SURVEY_SHOUT_DEMO["table_commitment"]
    assert P.g1_compressed(sp.commitments[1]).hex() == SURVEY_SHOUT_DEMO["index_commitment"]
    assert sp.z == SURVEY_SHOUT_DEMO["z"] and [P.g1_compressed(x).hex() for x in sp.opening_proofs] == SURVEY_SHOUT_DEMO["opening_proofs"]
    out["shout_demo"] = dict(log_size=3, entries=[i * i for i in range(8)], lookups=[3, 5, 0, 7], proof_hex=sp.to_bytes().hex(), z=str(sp.z))

    # README quick start: setup_params(8) - only the first few powers are needed
    pp8 = P.setup_params(8, max_powers=8)
    ops = [("W", 0, 42), ("W", 1, 100), ("R", 0, 42)]
    pr = P.twist_prove(pp8, ops)
    assert P.g1_compressed(pr.commitments[0]).hex() == SURVEY_README_TWIST["address_commitment"]
    assert P.g1_compressed(pr.commitments[1]).hex() == SURVEY_README_TWIST["value_commitment"]
    assert pr.z == SURVEY_README_TWIST["z"] and [P.g1_compressed(x).hex() for x in pr.opening_proofs] == SURVEY_README_TWIST["opening_proofs"]
    out["twist_readme"] = dict(log_size=8, ops=[[k, a, v] for k, a, v in ops], proof_hex=pr.to_bytes().hex(), z=str(pr.z))
    sp = P.shout_prove(pp8, [1, 4, 9], [1])
    assert P.g1_compressed(sp.commitments[0]).hex() == SURVEY_README_SHOUT["table_commitment"]
    assert P.g1_compressed(sp.commitments[1]).hex() == SURVEY_README_SHOUT["index_commitment"]
    out["shout_readme"] = dict(log_size=8, entries=[1, 4, 9], lookups=[1], proof_hex=sp.to_bytes().hex())
    # empty trace / no lookups (twist_tests.rs:88-99, shout_tests.rs:100-118)
    out["twist_empty"] = dict(log_size=3, ops=[], proof_hex=P.twist_prove(pp3, []).to_bytes().hex())
    out["shout_no_lookups"] = dict(log_size=3, entries=[1, 2, 3, 4], lookups=[], proof_hex=P.shout_prove(pp3, [1, 2, 3, 4], []).to_bytes().hex())

    # Appendix C.3 product sum-check
    A = list(range(1, 9)); B = [3, 1, 4, 1, 5, 9, 2, 6]
    tr = P.Transcript()
    amle = P.MultilinearExtension.from_evaluations(A); bmle = P.MultilinearExtension.from_evaluations(B)
    rps, fe = P.sumcheck_prove(3, 162, lambda v: amle.evaluate(v) * bmle.evaluate(v) % P.R_MOD, tr)
    assert rps[0] == SURVEY_C3["round0"] and fe == SURVEY_C3["final_evaluation"]
    tr2 = P.Transcript()
    rps2, fe2, _ = P.sumcheck_prove_product_tables([A, B], 162, tr2)
    assert rps2 == rps and fe2 == fe
    ok, ch = P.sumcheck_verify(3, 162, rps, fe, P.Transcript())
    assert ok and ch == SURVEY_C3["challenges"]
    out["sumcheck_c3"] = dict(A=A, B=B, claimed_sum=162, round_polynomials=[[str(c) for c in rp] for rp in rps],
                              challenges=[str(c) for c in ch], final_evaluation=str(fe))
    with open(os.path.join(HERE, "appendix_c.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote appendix_c.json")


if __name__ == "__main__":
    main()
