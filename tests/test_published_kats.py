"""CPU: PUBLISHED known-answer vectors of the third-party primitives the reference leans on, checked against the
pure-Python restatement (oracle/pyref.py), which the C++ oracle and the product's host code are then compared with on many inputs.
Sources, all reproduced here from the publications (no network):
  * SipHash-2-4, Aumasson & Bernstein, "SipHash: a fast short-input PRF", Appendix A: key 00..0f, message 00..0e -> a129ca6149be45e5.
    It pins the round function / padding / finalisation structure; SipHash-1-3 is the same code with (c, d) = (1, 3).
  * SipHash-1-3, Rust library/core/tests/hash/sip.rs `test_siphash_1_3`: key 00..0f, messages 00..(i-1), first three vectors.
  * Rust `DefaultHasher::new().finish()` (SipHash-1-3, keys 0, 0, empty input) = d1fba762150c532c.
  * ChaCha20: RFC 7539 section 2.3.2 (key 00..1f, counter 1, nonce 00:00:00:09:00:00:00:4a:00:00:00:00) and Appendix A.1 vectors 1 and 2
    (zero key and nonce, counters 0 and 1).  rand_chacha's ChaCha20Rng is this block function with a 64-bit counter / 64-bit stream id.
  * BN254 G1: generator (1, 2); 2G from EIP-196 (alt_bn128 addition tests); the curve equation y^2 = x^3 + 3.
What no publication fixes - ark-ff's Fp::rand masking and rejection, BlockRng's word order for next_u64, ark-serialize's flag bits - is
restated from the crates' documented behaviour and stays UNPINNED until tests/golden/from_arkworks.json exists (see the last test)."""
import json
import os

import numpy as np
import pytest

from conftest import seed_bytes

HERE = os.path.dirname(os.path.abspath(__file__))
M64 = (1 << 64) - 1


def _siphash(c, d, data, k0, k1):
    """generic SipHash-c-d, written independently of oracle/pyref.py"""
    def rotl(v, s):
        return ((v << s) | (v >> (64 - s))) & M64
    v = [k0 ^ 0x736F6D6570736575, k1 ^ 0x646F72616E646F6D, k0 ^ 0x6C7967656E657261, k1 ^ 0x7465646279746573]

    def rnd():
        v[0] = (v[0] + v[1]) & M64; v[1] = rotl(v[1], 13) ^ v[0]; v[0] = rotl(v[0], 32)
        v[2] = (v[2] + v[3]) & M64; v[3] = rotl(v[3], 16) ^ v[2]
        v[0] = (v[0] + v[3]) & M64; v[3] = rotl(v[3], 21) ^ v[0]
        v[2] = (v[2] + v[1]) & M64; v[1] = rotl(v[1], 17) ^ v[2]; v[2] = rotl(v[2], 32)
    n = len(data)
    words = [int.from_bytes(data[i:i + 8], "little") for i in range(0, n - n % 8, 8)]
    words.append(((n & 0xFF) << 56) | int.from_bytes(data[n - n % 8:], "little"))
    for m in words:
        v[3] ^= m
        for _ in range(c):
            rnd()
        v[0] ^= m
    v[2] ^= 0xFF
    for _ in range(d):
        rnd()
    return v[0] ^ v[1] ^ v[2] ^ v[3]


K0 = int.from_bytes(bytes(range(8)), "little"); K1 = int.from_bytes(bytes(range(8, 16)), "little")


def test_siphash_published_vectors_and_all_implementations(oracle, tsgpu):
    import pyref
    assert _siphash(2, 4, bytes(range(15)), K0, K1) == 0xA129CA6149BE45E5                    # SipHash paper, Appendix A
    rust_1_3 = [0xABAC0158050FC4DC, 0xC9F49BF37D57CA93, 0x82CB9B024DC7D44D]                 # Rust core tests, test_siphash_1_3, rows 0..2
    for i, want in enumerate(rust_1_3):
        assert _siphash(1, 3, bytes(range(i)), K0, K1) == want == pyref.siphash13(bytes(range(i)), K0, K1)
    assert _siphash(1, 3, b"", 0, 0) == 0xD1FBA762150C532C == oracle.siphash13(b"")         # DefaultHasher::new().finish()
    rng = np.random.default_rng(13)
    for n in list(range(0, 70)) + [127, 128, 129, 1000, 4097]:
        d = rng.bytes(n)
        assert oracle.siphash13(d) == pyref.siphash13(d) == _siphash(1, 3, d, 0, 0)
    # the product's own SipHash (host/transcript.hpp) is reached through a Transcript: challenge = ChaCha20Rng(seed = hash(len || state) LE x 4)
    for k in (0, 1, 5):
        t = tsgpu.Transcript(); o = oracle.Transcript(); p = pyref.Transcript()
        for j in range(k):
            x = oracle.fr_from_ints([j * 7 + 1])[0]
            t.append_field_element(b"label%d" % j, x); o.append_field_element(b"label%d" % j, x); p.append_field_element(b"label%d" % j, j * 7 + 1)
        c = t.challenge_field_element(b"c")
        assert (c == o.challenge_field_element(b"c")).all() and oracle.fr_to_ints(c.reshape(1, 4))[0] == p.challenge_field_element(b"c")


def test_chacha20_rfc7539_blocks_and_the_rng_word_stream(oracle, tsgpu):
    import pyref
    key = [int.from_bytes(bytes(range(4 * i, 4 * i + 4)), "little") for i in range(8)]
    blk = pyref._chacha_block(key, 1 | (0x09000000 << 32), 0x4A000000)                       # RFC 7539 2.3.2 in the 64 / 64-bit counter / stream layout
    assert b"".join(w.to_bytes(4, "little") for w in blk).hex() == (
        "10f1e7e4d13b5915500fdd1fa32071c4c7d1f4c733c068030422aa9ac3d46c4ed2826446079faa0914c2d705d98b02a2b5129cd1de164eb9cbd083e8a2503c4e")
    a1 = ["76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586",
          "9f07e7be5551387a98ba977c732d080dcb0f29a048e3656912c6533e32ee7aed29b721769ce64e43d57133b074d839d531ed1f28510afb45ace10a1f4b794d6f"]
    for ctr, want in enumerate(a1):                                                           # RFC 7539 A.1 test vectors 1 and 2
        assert b"".join(w.to_bytes(4, "little") for w in pyref._chacha_block([0] * 8, ctr)).hex() == want
    # ChaCha20Rng::from_seed([0; 32]): the keystream in order, next_u64 = two consecutive words, low word first (rand_core BlockRng)
    stream = bytes.fromhex(a1[0] + a1[1])
    want_u64 = [int.from_bytes(stream[8 * i:8 * i + 8], "little") for i in range(16)]
    assert [int(x) for x in oracle.chacha_u64(bytes(32), 16)] == want_u64 == [int(x) for x in tsgpu.chacha20_u64(bytes(32), 16)]
    # 300 draws cross the 64-word buffer four times; python, oracle and product agree
    rng = pyref.ChaCha20Rng(seed_bytes(3))
    py = [rng.next_u64() for _ in range(300)]
    assert py == [int(x) for x in oracle.chacha_u64(seed_bytes(3), 300)] == [int(x) for x in tsgpu.chacha20_u64(seed_bytes(3), 300)]
    # BlockRng::next_u64 at an ODD word index, including the straddle at word 63 -> 0 of the next buffer (unreachable from the reference, which
    # only ever draws u64s from a fresh generator, but part of the restated behaviour): one next_u32 first, then 40 next_u64
    rng = pyref.ChaCha20Rng(bytes(32))
    first = rng.next_u32()
    ks = b"".join(w.to_bytes(4, "little") for c in range(6) for w in pyref._chacha_block([0] * 8, c))
    assert first == int.from_bytes(ks[:4], "little")
    assert [rng.next_u64() for _ in range(40)] == [int.from_bytes(ks[4 + 8 * i:12 + 8 * i], "little") for i in range(40)]


def test_bn254_g1_published_points_and_encodings(oracle, tsgpu):
    p = oracle.P_MOD
    G = oracle.g1_generator()
    assert oracle.g1_affine_canonical(G) == [(1, 2)]
    two_g = (0x030644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD3, 0x15ED738C0E0A7C92E7845F96B2AE9C0A68A6A449E3538FC7FF3EBF7A5A18A2C4)   # EIP-196
    assert (two_g[1] ** 2 - two_g[0] ** 3 - 3) % p == 0
    assert oracle.g1_affine_canonical(oracle.g1_add(G, G)) == [two_g] == oracle.g1_affine_canonical(oracle.g1_mul(G, oracle.fr_from_ints([2])[0]))
    # r G = identity (the group order is the Fr modulus): (r - 1) G = -G
    neg_g = oracle.g1_mul(G, oracle.fr_from_ints([oracle.R_MOD - 1])[0])
    assert oracle.g1_affine_canonical(neg_g) == [(1, p - 2)]
    # ark-serialize compressed G1 (SWFlags): x little-endian; bit 7 of the last byte = "y is the larger of (y, -y)", bit 6 = infinity
    assert oracle.g1_compress(G).hex() == "01" + "00" * 31 == tsgpu.g1_compress(G).hex()
    assert oracle.g1_compress(neg_g).hex() == "01" + "00" * 30 + "80" == tsgpu.g1_compress(neg_g).hex()
    ident = oracle.g1_add(G, neg_g)
    assert oracle.g1_compress(ident).hex() == "00" * 31 + "40" == tsgpu.g1_compress(ident).hex()
    x2 = two_g[0].to_bytes(32, "little")
    flag = 0x80 if two_g[1] > p - two_g[1] else 0
    assert oracle.g1_compress(oracle.g1_add(G, G)) == x2[:31] + bytes([x2[31] | flag])


# EIP-196 scalar multiplication on a point that is not the generator: go-ethereum's `bn256ScalarMul` precompile test "chfast1" (input point, 64-bit scalar, expected product)
CHFAST1_POINT = (0x2bd3e6d0f3b142924f5ca7b49ce5b9d54c4703d7ae5648e61d02268b1a0a9fb7, 0x21611ce0a6af85915e2f1d70300909ce2e49dfad4a4619c8390cae66cefdb204)
CHFAST1_SCALAR = 0x11138ce750fa15c2
CHFAST1_PRODUCT = (0x070a8d6a982153cae4be29d434e8faef8a47b274a053f5a4ee2a6c9c13c31e5c, 0x031b8ce914eba3a9ffb989f9cdd5b0f01943074bf4f0f315690ec3cec6981afc)


def test_bn254_g1_published_scalar_multiplication(oracle):
    """the published alt_bn128 product k P (EIP-196 test "chfast1") from the pure-Python restatement and from the C++ oracle's Jacobian double-and-add:
    pins the G1 group law and `point * scalar` (commitments.rs:173-177, utils.rs:94-102) on a foreign point"""
    import pyref
    assert pyref.g1_is_on_curve(CHFAST1_POINT) and pyref.g1_is_on_curve(CHFAST1_PRODUCT)
    assert pyref.g1_mul(CHFAST1_POINT, CHFAST1_SCALAR) == CHFAST1_PRODUCT
    P = np.concatenate([oracle.fq_from_ints([CHFAST1_POINT[0]]).reshape(-1), oracle.fq_from_ints([CHFAST1_POINT[1]]).reshape(-1), oracle.fq_from_ints([1]).reshape(-1)])
    got = oracle.g1_mul(P, oracle.fr_from_ints([CHFAST1_SCALAR])[0])
    assert oracle.g1_affine_canonical(got) == [CHFAST1_PRODUCT]
    # EIP-196 point addition on two foreign points (go-ethereum's `bn256Add` test "chfast1")
    A = (0x18b18acfb4c2c30276db5411368e7185b311dd124691610c5d3b74034e093dc9, 0x063c909c4720840cb5134cb9f59fa749755796819658d32efc0d288198f37266)
    B = (0x07c2b7f58a84bd6145f00c9c2bc0bb1a187f20ff2c92963a88019e7c6a014eed, 0x06614e20c147e940f2d70da3f74c9a17df361706a4485c742bd6788478fa17d7)
    C = (0x2243525c5efd4b9c3d3c45ac0ca3fe4dd85e830a4ce6b65fa1eeaee202839703, 0x301d1d33be6da8e509df21cc35964723180eed7532537db9ae5e7d48f195c915)
    assert pyref.g1_add(A, B) == C == pyref.g1_add(B, A)
    jac = lambda Q: np.concatenate([oracle.fq_from_ints([Q[0]]).reshape(-1), oracle.fq_from_ints([Q[1]]).reshape(-1), oracle.fq_from_ints([1]).reshape(-1)])
    assert oracle.g1_affine_canonical(oracle.g1_add(jac(A), jac(B))) == [C]
    assert oracle.g1_affine_canonical(oracle.msm_pippenger(oracle.g1_batch_to_affine(np.stack([jac(A), jac(B)])), oracle.fr_from_ints([1, 1]))) == [C]
    # and through the CPU Pippenger the GPU MSM is compared with: sum over three copies with scalars k - 5, 2, 3
    aff = oracle.g1_batch_to_affine(np.stack([P, P, P]))
    assert oracle.g1_affine_canonical(oracle.msm_pippenger(aff, oracle.fr_from_ints([CHFAST1_SCALAR - 5, 2, 3]))) == [CHFAST1_PRODUCT]


def test_fp_rand_masking_and_rejection_branch(oracle):
    """ark-ff 0.4.2 `Fp::rand` (documented behaviour, UNPINNED by any publication): four next_u64 -> the top 2 bits cleared -> accepted as the
    MONTGOMERY representation iff < r, else the next four words.  Some seed below must exercise a rejection."""
    import pyref
    rejected = 0
    for s in range(1, 9):
        seed = seed_bytes(s)
        words = [int(x) for x in oracle.chacha_u64(seed, 64)]
        got = oracle.limbs_to_ints(oracle.chacha_fr_rand(seed, 4))        # raw limbs = Montgomery form
        k = 0
        for i in range(4):
            while True:
                cand = (words[k] | (words[k + 1] << 64) | (words[k + 2] << 128) | ((words[k + 3] & (M64 >> 2)) << 192)); k += 4
                if cand < oracle.R_MOD:
                    break
                rejected += 1
            assert got[i] == cand
        rng = pyref.ChaCha20Rng(seed)
        assert [pyref.fr_rand_mont(rng) for _ in range(4)] == got
    assert rejected > 0


def test_vectors_from_real_arkworks_when_present(oracle):
    """`cargo run --release -- golden` in baseline/arkworks_bench writes tests/golden/from_arkworks.json from REAL arkworks / rand_chacha /
    Rust std.  When that file exists every vector in it must equal the repository's self-derived tests/golden/appendix_c.json (which
    tests/test_oracle_golden.py holds the oracle to).  Until someone with a Rust toolchain commits it, byte parity with the reference is
    'partial': restated and cross-checked, not pinned."""
    path = os.path.join(HERE, "golden", "from_arkworks.json")
    if not os.path.exists(path):
        pytest.skip("tests/golden/from_arkworks.json absent: no Rust toolchain in this image - parity with real arkworks stays unpinned")
    real = json.load(open(path))
    ours = json.load(open(os.path.join(HERE, "golden", "appendix_c.json")))
    checked = 0
    for key, val in real.items():
        if key == "provenance":
            continue
        assert key in ours, key
        if isinstance(val, dict):
            for sub, v in val.items():
                assert ours[key][sub] == v, (key, sub)
                checked += 1
        else:
            assert ours[key] == val, key
            checked += 1
    assert checked >= 20


def test_bn254_pairing_check_vector_from_the_alt_bn128_precompile_tests(tsgpu):
    """EIP-197 pairing check e(P1, Q1) e(P2, Q2) == 1 on points this repository did not produce: the 384-byte input of go-ethereum's `bn256Pairing` precompile test
    "jeff1" (expected output 1; the second G2 point is the standard generator).  Encoding there: big-endian 32-byte words, G1 = (x, y), G2 = (x.im, x.re, y.im, y.re).
    The product must be one as published, stop being one when a point is negated or swapped for another valid point, and off-curve / unreduced input must be refused.
    (A product check is invariant under powers of the pairing, so this pins bilinearity on foreign points, not the GT element itself.)"""
    import ctypes as C
    words = """1c76476f4def4bb94541d57ebba1193381ffa7aa76ada664dd31c16024c43f59 3034dd2920f673e204fee2811c678745fc819b55d3e9d294e45c9b03a76aef41
               209dd15ebff5d46c4bd888e51a93cf99a7329636c63514396b4a452003a35bf7 04bf11ca01483bfa8b34b43561848d28905960114c8ac04049af4b6315a41678
               2bb8324af6cfc93537a2ad1a445cfd0ca2a71acd7ac41fadbf933c2a51be344d 120a2a4cf30c1bf9845f20c6fe39e07ea2cce61f0c9bb048165fe5e4de877550
               111e129f1cf1097710d41c4ac70fcdfa5ba2023c6ff1cbeac322de49d1b6df7c 2032c61a830e3c17286de9462bf242fca2883585b93870a73853face6a6bf411
               198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2 1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed
               090689d0585ff075ec9e99ad690c3395bc4b313370b38ef355acdadcd122975b 12c85ea5db8c6deb4aab71808dcb408fe3d1e7690c43d37b4ce6cc0166fa7daa""".split()
    w = [int(x, 16) for x in words]
    P_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583

    def limbs(vals):
        return np.array([(v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for v in vals for k in range(4)], dtype=np.uint64)

    def check(g1_pts, g2_pts):
        g1 = limbs([c for pt in g1_pts for c in pt]); g2 = limbs([c for pt in g2_pts for c in pt])
        return tsgpu.lib().tsgpu_pairing_check_points(g1.ctypes.data_as(C.c_void_p), g2.ctypes.data_as(C.c_void_p), C.c_size_t(len(g1_pts)))

    p1, p2 = (w[0], w[1]), (w[6], w[7])
    q1, q2 = (w[3], w[2], w[5], w[4]), (w[9], w[8], w[11], w[10])          # (re, im) order of the hook
    assert check([p1, p2], [q1, q2]) == 1
    assert check([p2, p1], [q2, q1]) == 1
    assert check([p1], [q1]) == 0 and check([p2], [q2]) == 0              # each factor alone is not one
    assert check([(p1[0], P_MOD - p1[1]), p2], [q1, q2]) == 0             # -P1
    assert check([p1, p2], [q2, q2]) == 0                                  # another valid G2 point
    assert check([p1, p1], [q1, (q1[0], q1[1], P_MOD - q1[2], (P_MOD - q1[3]) % P_MOD)]) == 1    # e(P, Q) e(P, -Q) = 1
    assert check([(0, 0), p2], [q1, (0, 0, 0, 0)]) == 1                   # identities contribute one
    assert check([(p1[0], p1[1] + 1), p2], [q1, q2]) == -1                # off the curve
    assert check([(p1[0] + P_MOD, p1[1]), p2], [q1, q2]) == -1            # not reduced
    assert check([p1, p2], [(q1[0] + 1, q1[1], q1[2], q1[3]), q2]) == -1  # G2 point off the twist
