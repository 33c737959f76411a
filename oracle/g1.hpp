// oracle/g1.hpp - CPU restatement of ark-ec 0.4.2 short-Weierstrass G1 for BN254
// (y^2 = x^3 + 3 over Fq, generator (1,2), Jacobian coordinates; reference alias
// src/utils.rs:17 `G1Element = G1Projective`).  TEST INFRASTRUCTURE ONLY.
//
// ark-ec is not vendored under /root/reference (Cargo.lock: ark-ec 0.4.2, ark-bn254 0.4.0).
// Only the resulting group element matters for parity (any Jacobian representative
// normalises to the same affine point), so the standard a=0 formulas are used:
// dbl-2009-l, add-2007-bl, madd-2007-bl (EFD).
#pragma once
#include "ff.hpp"

namespace orc {

struct G1Affine {
    Fq x, y;
    bool inf;
};

struct G1 {   // Jacobian: (X/Z^2, Y/Z^3); identity <=> Z == 0.  96 bytes, same order as arkworks {x,y,z}.
    Fq X, Y, Z;

    static G1 identity() { G1 r; r.X = Fq::one(); r.Y = Fq::one(); r.Z = Fq::zero(); return r; }
    static G1 generator() { G1 r; r.X = Fq::from_u64(1); r.Y = Fq::from_u64(2); r.Z = Fq::one(); return r; }
    static G1 from_affine(const G1Affine& a) {
        if (a.inf) return identity();
        G1 r; r.X = a.x; r.Y = a.y; r.Z = Fq::one(); return r;
    }
    bool is_identity() const { return Z.is_zero(); }

    G1 dbl() const {
        if (is_identity()) return *this;
        Fq A = X.sqr(), B = Y.sqr(), C = B.sqr();
        Fq D = ((X + B).sqr() - A - C).dbl();
        Fq E = A.dbl() + A, F = E.sqr();
        G1 r;
        r.X = F - D.dbl();
        r.Y = E * (D - r.X) - C.dbl().dbl().dbl();
        r.Z = (Y * Z).dbl();
        return r;
    }
    G1 add(const G1& o) const {
        if (is_identity()) return o;
        if (o.is_identity()) return *this;
        Fq Z1Z1 = Z.sqr(), Z2Z2 = o.Z.sqr();
        Fq U1 = X * Z2Z2, U2 = o.X * Z1Z1;
        Fq S1 = Y * o.Z * Z2Z2, S2 = o.Y * Z * Z1Z1;
        if (U1 == U2) {
            if (S1 == S2) return dbl();
            return identity();
        }
        Fq H = U2 - U1, I = H.dbl().sqr(), J = H * I, rr = (S2 - S1).dbl(), V = U1 * I;
        G1 r;
        r.X = rr.sqr() - J - V.dbl();
        r.Y = rr * (V - r.X) - (S1 * J).dbl();
        r.Z = ((Z + o.Z).sqr() - Z1Z1 - Z2Z2) * H;
        return r;
    }
    G1 add_mixed(const G1Affine& o) const {
        if (o.inf) return *this;
        if (is_identity()) return from_affine(o);
        Fq Z1Z1 = Z.sqr();
        Fq U2 = o.x * Z1Z1, S2 = o.y * Z * Z1Z1;
        if (X == U2) {
            if (Y == S2) return dbl();
            return identity();
        }
        Fq H = U2 - X, HH = H.sqr(), I = HH.dbl().dbl(), J = H * I, rr = (S2 - Y).dbl(), V = X * I;
        G1 r;
        r.X = rr.sqr() - J - V.dbl();
        r.Y = rr * (V - r.X) - (Y * J).dbl();
        r.Z = (Z + H).sqr() - Z1Z1 - HH;
        return r;
    }
    G1 neg() const { G1 r = *this; r.Y = Y.neg(); return r; }

    // ark-ec `Mul<Fr>`: MSB-first double-and-add over the canonical scalar bits.
    G1 mul(const Fr& k) const {
        uint64_t e[4]; k.to_canonical_limbs(e);
        G1 acc = identity();
        bool started = false;
        for (int i = 255; i >= 0; --i) {
            if (started) acc = acc.dbl();
            if ((e[i / 64] >> (i % 64)) & 1) { acc = acc.add(*this); started = true; }
        }
        return acc;
    }
    G1Affine to_affine() const {
        G1Affine a;
        if (is_identity()) { a.x = Fq::zero(); a.y = Fq::zero(); a.inf = true; return a; }
        Fq zi = Z.inverse(), zi2 = zi.sqr();
        a.x = X * zi2; a.y = Y * zi2 * zi; a.inf = false;
        return a;
    }
    bool equals(const G1& o) const {   // cross-multiplied projective equality
        if (is_identity() || o.is_identity()) return is_identity() && o.is_identity();
        Fq Z1Z1 = Z.sqr(), Z2Z2 = o.Z.sqr();
        return X * Z2Z2 == o.X * Z1Z1 && Y * o.Z * Z2Z2 == o.Y * Z * Z1Z1;
    }
};

inline bool fq_lex_greater_than_neg(const Fq& y) {
    // ark-serialize SWFlags::from_y_coordinate: negative iff y > -y (canonical integers)
    uint64_t a[4], b[4];
    y.to_canonical_limbs(a);
    y.neg().to_canonical_limbs(b);
    for (int i = 3; i >= 0; --i) {
        if (a[i] > b[i]) return true;
        if (a[i] < b[i]) return false;
    }
    return false;
}

// ark-serialize 0.4.2 compressed G1: x (32 B LE canonical) | flags in top bits of byte 31
inline void g1_compress(const G1& p, uint8_t out[32]) {
    G1Affine a = p.to_affine();
    if (a.inf) { memset(out, 0, 32); out[31] |= 0x40; return; }
    a.x.to_bytes_le(out);
    if (fq_lex_greater_than_neg(a.y)) out[31] |= 0x80;
}

// KZGCommitmentValue::hash (src/commitments.rs:73-84): affine x canonical bytes -> Fr mod r
inline Fr g1_hash(const G1& p) {
    G1Affine a = p.to_affine();
    if (a.inf) return Fr::zero();
    uint64_t c[4]; a.x.to_canonical_limbs(c);
    return Fr::from_canonical_limbs(c);
}

}  // namespace orc
