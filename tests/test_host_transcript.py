"""CPU: the product's host-side Transcript / SumCheck::verify (C++ in host/, through the C ABI) against the
oracle and the pinned vectors.  No GPU needed - these entry points never touch the device."""
import json
import os

import numpy as np
import pytest

from conftest import seed_bytes

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_transcript_known_answer(tsgpu, oracle):
    """src/utils.rs:287-296 test_transcript inputs; expected value from SURVEY.md Appendix C.1."""
    t = tsgpu.Transcript(bytes([42]) * 32)
    t.append_field_element(b"test", oracle.fr_from_ints([123])[0])
    c = t.challenge_field_element(b"challenge")
    assert oracle.fr_to_ints(c)[0] == 13648926573440158680322210633940909009220968087751212041477676025471912345605


def test_transcript_matches_oracle_on_random_sessions(tsgpu, oracle):
    rng = np.random.default_rng(7)
    for trial in range(20):
        a, b = tsgpu.Transcript(), oracle.Transcript()
        for step in range(int(rng.integers(1, 12))):
            label = bytes(rng.integers(97, 123, size=int(rng.integers(0, 40)), dtype=np.uint8))
            if rng.random() < 0.6:
                n = int(rng.integers(0, 6))
                xs = oracle.chacha_fr_rand(seed_bytes(trial * 16 + step), n).reshape(n, 4)
                a.append_field_elements(label, xs); b.append_field_elements(label, xs)
            else:
                assert (a.challenge_field_element(label) == b.challenge_field_element(label)).all()
        assert (a.challenge_field_elements(b"opening_challenges", 3) == b.challenge_field_elements(b"opening_challenges", 3)).all()


def test_sumcheck_verify_host(tsgpu, oracle):
    """verify() replays the transcript and accepts an honest proof / rejects tampering (sumcheck.rs:113-153)."""
    A = oracle.chacha_fr_rand(seed_bytes(1), 32); B = oracle.chacha_fr_rand(seed_bytes(2), 32)
    ai, bi = oracle.fr_to_ints(A), oracle.fr_to_ints(B)
    claimed = oracle.fr_from_ints([sum(x * y for x, y in zip(ai, bi)) % oracle.R_MOD])[0]
    ref = oracle.sumcheck_prove_product([A, B], claimed, mode="closure")
    proof = tsgpu.SumCheckProof(ref["round_polynomials"], ref["final_evaluation"])
    ok, ch = tsgpu.SumCheck(5, claimed).verify(proof, tsgpu.Transcript())
    assert ok and (ch == ref["challenges"]).all()
    bad = tsgpu.SumCheckProof(ref["round_polynomials"].copy(), ref["final_evaluation"])
    bad.round_polynomials[2, 1, 0] ^= 1
    ok, _ = tsgpu.SumCheck(5, claimed).verify(bad, tsgpu.Transcript())
    assert not ok
    with pytest.raises(tsgpu.TwistAndShoutError) as e:
        tsgpu.SumCheck(4, claimed).verify(proof, tsgpu.Transcript())
    assert e.value.variant == "SumCheck"


def test_g1_host_helpers_match_oracle(tsgpu, oracle):
    """hash (commitments.rs:73-84), compressed bytes and equality of G1 points - CPU-only entry points."""
    g = oracle.g1_generator()
    pts = [g, oracle.g1_add(g, g), oracle.g1_mul(g, oracle.fr_from_ints([123456789])[0]),
           oracle.g1_mul(g, oracle.fr_from_ints([oracle.R_MOD - 1])[0]), oracle.g1_mul(g, oracle.fr_from_ints([0])[0])]
    for p in pts:
        assert tsgpu.g1_compress(p) == oracle.g1_compress(p)
        assert (tsgpu.g1_hash(p) == oracle.g1_hash(p)).all()
    assert tsgpu.g1_compress(g).hex() == "01" + "00" * 31            # 1*G = (1, 2)
    assert tsgpu.g1_equal(pts[1], oracle.g1_mul(g, oracle.fr_from_ints([2])[0]))
    assert not tsgpu.g1_equal(pts[1], pts[2])


def test_transcript_long_session_covers_rejections_and_growing_state(tsgpu, oracle):
    """500 challenges over a state that grows to ~36 KB: Fp::rand rejects a draw with probability ~1/4, so runs of 3+ rejections (ChaCha blocks
    generated on demand beyond the first, host/transcript.hpp) and every state length modulo 8 (the word-wise SipHash tail) occur."""
    a, b = tsgpu.Transcript(), oracle.Transcript()
    xs = oracle.chacha_fr_rand(seed_bytes(9), 4).reshape(4, 4)
    for i in range(500):
        label = b"r" * (i % 11)
        a.append_field_elements(label, xs[: i % 5]); b.append_field_elements(label, xs[: i % 5])
        assert (a.challenge_field_element(b"c%d" % i) == b.challenge_field_element(b"c%d" % i)).all(), i
    assert a.state_len > 30000


def test_less_than_on_field_elements_host(tsgpu, oracle):
    """LessThanPolynomial::evaluate_at_field_elements (src/polynomials.rs:213-220, :266-283): the low num_vars bits of the canonical integers,
    first differing bit from bit 0 decides; host-only entry point"""
    rng = np.random.default_rng(3)
    for nv in (1, 3, 8, 64, 70, 254):
        lt = tsgpu.LessThanPolynomial.new(nv)
        for _ in range(40):
            x, y = int(rng.integers(0, 1 << 62)) << int(rng.integers(0, 190)), int(rng.integers(0, 1 << 62)) << int(rng.integers(0, 190))
            x %= oracle.R_MOD; y %= oracle.R_MOD
            want = 0
            for i in range(nv):
                bx, by = (x >> i) & 1, (y >> i) & 1
                if bx != by:
                    want = 1 if by else 0
                    break
            got = lt.evaluate_at_field_elements(oracle.fr_from_ints([x])[0], oracle.fr_from_ints([y])[0])
            assert oracle.fr_to_ints(got) == [want]


def test_statement_digest_matches_hashlib_blake2b_tree(tsgpu, oracle):
    """host/statement_digest.hpp (BLAKE2b-256, RFC 7693, two-level tree over 2^20-byte chunks) against hashlib: empty, sub-block, block-boundary,
    multi-chunk and ragged segments; and the RFC's "abc" vector through a one-segment leaf"""
    import hashlib
    rng = np.random.default_rng(11)

    def want(domain, header, segs):
        root = hashlib.blake2b(digest_size=32)
        root.update(domain[:16].ljust(16, b"\0") + len(header).to_bytes(8, "little") + b"".join(int(h).to_bytes(8, "little") for h in header))
        root.update(len(segs).to_bytes(8, "little"))
        for s in segs:
            root.update(len(s).to_bytes(8, "little"))
            for off in range(0, len(s), 1 << 20):
                root.update(hashlib.blake2b(s[off:off + (1 << 20)], digest_size=32).digest())
        return root.digest()

    cases = [(b"x", [], []), (b"shout_read_check", [3, 4], [b"", b"abc", bytes(127), bytes(128), bytes(129)]),
             (b"a-domain-longer-than-16", [1 << 63], [rng.bytes((1 << 20) + 5), rng.bytes(1 << 20), rng.bytes(3 * (1 << 20) - 1)])]
    for domain, header, segs in cases:
        assert tsgpu.statement_digest(domain, header, segs) == want(domain, header, segs)
    # the oracle's field-element form of the same digest
    d = tsgpu.statement_digest(b"shout_read_check", [3, 4], [b"abc"])
    got = oracle.fr_to_ints(oracle.statement_digest_elements(b"shout_read_check", [3, 4], [b"abc"]))
    assert got == [int.from_bytes(d[:16], "little"), int.from_bytes(d[16:], "little")]
    assert hashlib.blake2b(b"abc", digest_size=64).hexdigest().startswith("ba80a53f981c4d0d")          # RFC 7693 appendix A anchors hashlib itself


def test_chacha20_u64_matches_oracle_stream(tsgpu, oracle):
    """tsgpu_chacha20_u64 == the oracle's ChaCha20Rng::next_u64 stream, across several 64-word buffer refills"""
    for seed in (bytes(32), bytes([2]) * 32, bytes(range(32))):
        assert (tsgpu.chacha20_u64(seed, 1000) == oracle.chacha_u64(seed, 1000)).all()


def test_chacha20_fr_then_u64_matches_oracle(tsgpu, oracle):
    """Fr::rand draws (incl. the rejection branch: ~1 in 4 candidates is >= r after masking to 254 bits) followed by u64 draws from the same generator"""
    for seed in (bytes([4]) * 32, bytes([42]) * 32, bytes([5]) * 32):
        f, u = tsgpu.chacha20_fr_then_u64(seed, 300, 50)
        of, ou = oracle.chacha_fr_then_u64(seed, 300, 50)
        assert (f == of).all() and (u == ou).all()
    tau, _ = oracle.setup_scalars()
    assert (tsgpu.chacha20_fr_then_u64(bytes([42]) * 32, 1)[0][0] == tau).all()         # src/utils.rs:81-84
