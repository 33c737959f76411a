"""CPU: the BN254 pairing behind KZGCommitment::verify / batch_verify (src/commitments.rs:201-301), host-only.
No arkworks is available to compare GT elements with, so the pairing is pinned by its defining properties
(bilinearity, non-degeneracy, G2 generator of order r on the twist) and by the verification equation on openings
produced by the CPU oracle: honest openings verify, any tampering is rejected."""
import ctypes as C

import numpy as np
import pytest

from conftest import seed_bytes


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_g2_generator_and_bilinearity(tsgpu, oracle):
    L = tsgpu.lib()
    assert L.tsgpu_g2_generator_checks() == 1
    R = oracle.R_MOD
    for a, b in ((5, 7), (123456789, R - 2)):
        A = oracle.fr_from_ints([a, R - (a * b) % R]); B = oracle.fr_from_ints([b, 1])
        assert L.tsgpu_pairing_product_of_generators_is_one(_p(A), _p(B), C.c_size_t(2)) == 1      # e(aG, bH) e(-abG, H) = 1
        A = oracle.fr_from_ints([a, R - (a * b + 1) % R])
        assert L.tsgpu_pairing_product_of_generators_is_one(_p(A), _p(B), C.c_size_t(2)) == 0
    one = oracle.fr_from_ints([1])
    assert L.tsgpu_pairing_product_of_generators_is_one(_p(one), _p(one), C.c_size_t(1)) == 0           # e(G, H) != 1
    zero = oracle.fr_from_ints([0])
    assert L.tsgpu_pairing_product_of_generators_is_one(_p(zero), _p(one), C.c_size_t(1)) == 1          # e(O, H) = 1


def test_kzg_verify_on_oracle_openings(tsgpu, oracle):
    """src/commitments.rs:495-541: f(5) = 86 verifies, a wrong value does not; plus a random degree-63 polynomial"""
    vp = tsgpu.HostVerifierParams(4)
    pw = oracle.setup_g1_powers(65, fast=True)
    for poly, z in ((oracle.fr_from_ints([1, 2, 3]), oracle.fr_from_ints([5])[0]),
                    (oracle.chacha_fr_rand(seed_bytes(7), 64), oracle.chacha_fr_rand(seed_bytes(8), 1)[0])):
        Cm = oracle.kzg_commit(pw, poly)
        v, q = oracle.kzg_value_quotient(poly, z)
        pi = oracle.kzg_commit(pw, q)
        assert tsgpu.kzg_verify(vp, Cm, z, v, pi)
        assert oracle.kzg_check_trapdoor(Cm, z, v, pi)
        bad_v = oracle.field_binop("fr", "add", v.reshape(1, 4), oracle.fr_from_ints([1]))[0]
        assert not tsgpu.kzg_verify(vp, Cm, z, bad_v, pi)
        assert not tsgpu.kzg_verify(vp, Cm, bad_v, v, pi)                       # wrong point
        assert not tsgpu.kzg_verify(vp, oracle.g1_add(Cm, oracle.g1_generator()), z, v, pi)   # wrong commitment
    # constant polynomial: empty quotient, identity proof (commitments.rs:353-355)
    c7 = oracle.fr_from_ints([7])
    Cm = oracle.kzg_commit(pw, c7)
    ident = oracle.kzg_commit(pw, np.empty((0, 4), dtype=np.uint64))
    assert tsgpu.kzg_verify(vp, Cm, oracle.fr_from_ints([5])[0], c7[0], ident)


def test_kzg_batch_verify_mirrors_reference_formula(tsgpu, oracle):
    """KZGCommitment::batch_verify (src/commitments.rs:230-301) pairs sum_i gamma_i pi_i with sum_j gamma_j ([tau]_2 - [z_j]_2): gamma enters
    the right-hand side twice (and cross terms appear for n > 1), so the reference formula rejects every non-empty batch, honest or
    not; the reference never calls or tests it.  Parity means reproducing exactly that: empty batch -> true, anything else -> false,
    mismatched lengths -> Err(Commitment)."""
    vp = tsgpu.HostVerifierParams(4)
    pw = oracle.setup_g1_powers(33, fast=True)
    Cs, zs, vs, pis = [], [], [], []
    for k in range(3):
        poly = oracle.chacha_fr_rand(seed_bytes(20 + k), 16 + k)
        z = oracle.chacha_fr_rand(seed_bytes(30 + k), 1)[0]
        v, q = oracle.kzg_value_quotient(poly, z)
        Cs.append(oracle.kzg_commit(pw, poly)); zs.append(z); vs.append(v); pis.append(oracle.kzg_commit(pw, q))
        assert tsgpu.kzg_verify(vp, Cs[-1], z, v, pis[-1])
    assert not tsgpu.kzg_batch_verify(vp, Cs[:1], zs[:1], vs[:1], pis[:1])
    assert not tsgpu.kzg_batch_verify(vp, Cs[:1], zs[:1], [oracle.fr_from_ints([99])[0]], pis[:1])
    assert not tsgpu.kzg_batch_verify(vp, Cs, zs, vs, pis)
    e0 = np.empty((0, 12), dtype=np.uint64); f0 = np.empty((0, 4), dtype=np.uint64)
    assert tsgpu.kzg_batch_verify(vp, e0, f0, f0, e0)
    with pytest.raises(tsgpu.TwistAndShoutError) as e:
        tsgpu.kzg_batch_verify(vp, Cs, zs[:2], vs, pis)
    assert e.value.variant == "Commitment"


def test_pairing_fast_paths_agree_with_the_plain_ones(tsgpu):
    """split final exponentiation == f^((p^12 - 1) / r), Fq12 inverse / symmetric square / Frobenius^6 == conjugation, Jacobian G2 multiplication"""
    assert tsgpu.lib().tsgpu_pairing_self_check() == 1


def test_kzg_verify_rejects_bytes_that_are_not_points_of_g1(tsgpu, oracle):
    """The C ABI takes raw Jacobian coordinates; arkworks' G1Projective can only ever hold curve points.  An off-curve commitment or proof, or a
    coordinate that is not reduced mod p, makes verify return false instead of feeding the pairing undefined input."""
    vp = tsgpu.HostVerifierParams(4)
    pw = oracle.setup_g1_powers(9, fast=True)
    poly = oracle.fr_from_ints([1, 2, 3]); z = oracle.fr_from_ints([5])[0]
    Cm = oracle.kzg_commit(pw, poly)
    v, q = oracle.kzg_value_quotient(poly, z)
    pi = oracle.kzg_commit(pw, q)
    assert tsgpu.kzg_verify(vp, Cm, z, v, pi)
    off = Cm.copy(); off[0] ^= np.uint64(1)                                   # X changed: Y^2 != X^3 + 3 Z^6
    assert not tsgpu.kzg_verify(vp, off, z, v, pi) and not tsgpu.kzg_verify(vp, Cm, z, v, off)
    big = Cm.copy(); big[0:4] = oracle.int_to_limbs((1 << 256) - 1)           # X >= p: not a reduced field element
    assert not tsgpu.kzg_verify(vp, big, z, v, pi)
    # another Jacobian representative of the SAME point (X l^2, Y l^3, Z l) still verifies
    lam = 0x1234567
    x, y, zz = (oracle.fq_to_ints(Cm[4 * i:4 * i + 4].reshape(1, 4))[0] for i in range(3))
    rep = np.concatenate([oracle.fq_from_ints([x * lam * lam, y * lam ** 3, zz * lam]).reshape(-1)])
    assert oracle.g1_equal(rep, Cm) and tsgpu.kzg_verify(vp, rep, z, v, pi)
