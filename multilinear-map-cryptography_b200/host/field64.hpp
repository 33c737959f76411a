// host/field64.hpp - native 4 x 64-bit Montgomery arithmetic for the few field operations that stay on the
// CPU in the product: combining MSM window sums, affine normalisation for KZGCommitmentValue::hash
// (src/commitments.rs:73-84) and ark-serialize style compression of the proof points.
// Same element layout as the reference (ark-ff BigInt<4>, Montgomery, R = 2^256).
#pragma once
#include <cstdint>
#include <cstring>
#include "../csrc/fp.cuh"

namespace tsg {
namespace host {

typedef unsigned __int128 u128_t;

template <class P>
struct F64 {
    uint64_t l[4];

    static uint64_t modl(int i) { return (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32); }
    static uint64_t inv64() {
        // -p^-1 mod 2^64 by Newton iteration from the 32-bit constant
        uint64_t p0 = modl(0), x = (uint64_t)(0u - P::INV);   // p^-1 mod 2^32
        x *= 2 - p0 * x;                                       // now mod 2^64
        return 0 - x;
    }
    static F64 zero() { F64 r; r.l[0] = r.l[1] = r.l[2] = r.l[3] = 0; return r; }
    static F64 one() { F64 r; for (int i = 0; i < 4; ++i) r.l[i] = (uint64_t)P::one(2 * i) | ((uint64_t)P::one(2 * i + 1) << 32); return r; }
    static F64 r2() { F64 r; for (int i = 0; i < 4; ++i) r.l[i] = (uint64_t)P::r2(2 * i) | ((uint64_t)P::r2(2 * i + 1) << 32); return r; }
    static F64 from_raw(const void* p) { F64 r; memcpy(r.l, p, 32); return r; }
    static F64 from_u64(uint64_t v) { F64 r = zero(); r.l[0] = v; return r * r2(); }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
    bool operator==(const F64& o) const { return !memcmp(l, o.l, 32); }
    bool operator!=(const F64& o) const { return !(*this == o); }

    static bool geq_mod(const uint64_t* a) {
        for (int i = 3; i >= 0; --i) { uint64_t m = modl(i); if (a[i] > m) return true; if (a[i] < m) return false; }
        return true;
    }
    static void sub_mod(uint64_t* a) {
        u128_t br = 0;
        for (int i = 0; i < 4; ++i) { u128_t d = (u128_t)a[i] - modl(i) - br; a[i] = (uint64_t)d; br = (d >> 64) & 1; }
    }
    F64 operator+(const F64& o) const {
        F64 r; u128_t c = 0;
        for (int i = 0; i < 4; ++i) { c += (u128_t)l[i] + o.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
        if (geq_mod(r.l)) sub_mod(r.l);
        return r;
    }
    F64 operator-(const F64& o) const {
        F64 r; u128_t br = 0;
        for (int i = 0; i < 4; ++i) { u128_t d = (u128_t)l[i] - o.l[i] - br; r.l[i] = (uint64_t)d; br = (d >> 64) & 1; }
        if (br) { u128_t c = 0; for (int i = 0; i < 4; ++i) { c += (u128_t)r.l[i] + modl(i); r.l[i] = (uint64_t)c; c >>= 64; } }
        return r;
    }
    F64 neg() const { return is_zero() ? *this : zero() - *this; }
    F64 dbl() const { return *this + *this; }
    F64 operator*(const F64& o) const {
        static const uint64_t ninv = inv64();
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            u128_t c = 0;
            for (int j = 0; j < 4; ++j) { c += (u128_t)l[j] * o.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
            c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * ninv;
            c = (u128_t)m * modl(0) + t[0]; c >>= 64;
            for (int j = 1; j < 4; ++j) { c += (u128_t)m * modl(j) + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
            c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
        }
        F64 r; memcpy(r.l, t, 32);
        if (t[4] || geq_mod(r.l)) sub_mod(r.l);
        return r;
    }
    F64 sqr() const { return *this * *this; }
    F64 inverse_fermat() const {   // a^(p - 2); 0 -> 0.  Kept as the cross-check of inverse() (tsgpu_pairing_self_check)
        uint64_t e[4] = {modl(0) - 2, modl(1), modl(2), modl(3)};
        F64 acc = one();
        for (int i = 255; i >= 0; --i) { acc = acc.sqr(); if ((e[i >> 6] >> (i & 63)) & 1) acc = acc * *this; }
        return acc;
    }
    // Binary extended Euclid on the Montgomery representative x = a R (as an integer below p): y = x^-1 mod p, then a^-1 R = y R^2 - two products by R^2.
    // ~4x faster than the Fermat power (a few hundred 256-bit shifts / subtractions instead of ~380 products); the host inverts on every commitment hash, every
    // barycentric opening and at every step of the pairing's affine G2 arithmetic.  0 -> 0.
    F64 inverse() const {
        uint64_t u[4], v[4], b[4] = {1, 0, 0, 0}, c[4] = {0, 0, 0, 0};
        memcpy(u, l, 32);
        while (geq_mod(u)) sub_mod(u);   // limbs that arrive over the C ABI need not be reduced (2^256 < 6 p: a few subtractions); a multiple of p has no inverse and would never terminate below
        if ((u[0] | u[1] | u[2] | u[3]) == 0) return zero();
        for (int i = 0; i < 4; ++i) v[i] = modl(i);
        auto is_one = [](const uint64_t* a) { return a[0] == 1 && (a[1] | a[2] | a[3]) == 0; };
        auto shr1 = [](uint64_t* a) { a[0] = (a[0] >> 1) | (a[1] << 63); a[1] = (a[1] >> 1) | (a[2] << 63); a[2] = (a[2] >> 1) | (a[3] << 63); a[3] >>= 1; };
        auto add_p = [](uint64_t* a) { u128_t cy = 0; for (int i = 0; i < 4; ++i) { cy += (u128_t)a[i] + modl(i); a[i] = (uint64_t)cy; cy >>= 64; } };   // a < p: a + p < 2^255
        auto sub = [](uint64_t* a, const uint64_t* o) { u128_t br = 0; for (int i = 0; i < 4; ++i) { u128_t d = (u128_t)a[i] - o[i] - br; a[i] = (uint64_t)d; br = (d >> 64) & 1; } return (bool)br; };
        auto geq = [](const uint64_t* a, const uint64_t* o) { for (int i = 3; i >= 0; --i) { if (a[i] > o[i]) return true; if (a[i] < o[i]) return false; } return true; };
        // invariants: b x == u, c x == v (mod p); u, v > 0; b, c in [0, p)
        while (!is_one(u) && !is_one(v)) {
            while (!(u[0] & 1)) { shr1(u); if (b[0] & 1) add_p(b); shr1(b); }
            while (!(v[0] & 1)) { shr1(v); if (c[0] & 1) add_p(c); shr1(c); }
            if (geq(u, v)) { sub(u, v); if (sub(b, c)) add_p(b); }
            else { sub(v, u); if (sub(c, b)) add_p(c); }
        }
        F64 y; memcpy(y.l, is_one(u) ? b : c, 32);
        return (y * r2()) * r2();
    }
    F64 from_mont() const { F64 o = zero(); o.l[0] = 1; return *this * o; }
};

typedef F64<FrP> Fr64;
typedef F64<FqP> Fq64;

// reference G1Projective (Jacobian, a = 0): dbl-2009-l / add-2007-bl
struct G1J {
    Fq64 x, y, z;
    static G1J identity() { G1J r; r.x = Fq64::one(); r.y = Fq64::one(); r.z = Fq64::zero(); return r; }
    static G1J generator() { G1J r; r.x = Fq64::from_u64(1); r.y = Fq64::from_u64(2); r.z = Fq64::one(); return r; }
    bool is_identity() const { return z.is_zero(); }
    G1J dbl() const {
        if (is_identity()) return *this;
        Fq64 A = x.sqr(), B = y.sqr(), C = B.sqr();
        Fq64 D = ((x + B).sqr() - A - C).dbl();
        Fq64 E = A.dbl() + A, F = E.sqr();
        G1J r;
        r.x = F - D.dbl();
        r.y = E * (D - r.x) - C.dbl().dbl().dbl();
        r.z = (y * z).dbl();
        return r;
    }
    G1J add(const G1J& o) const {
        if (is_identity()) return o;
        if (o.is_identity()) return *this;
        Fq64 Z1Z1 = z.sqr(), Z2Z2 = o.z.sqr();
        Fq64 U1 = x * Z2Z2, U2 = o.x * Z1Z1, S1 = y * o.z * Z2Z2, S2 = o.y * z * Z1Z1;
        if (U1 == U2) return S1 == S2 ? dbl() : identity();
        Fq64 H = U2 - U1, I = H.dbl().sqr(), J = H * I, rr = (S2 - S1).dbl(), V = U1 * I;
        G1J r;
        r.x = rr.sqr() - J - V.dbl();
        r.y = rr * (V - r.x) - (S1 * J).dbl();
        r.z = ((z + o.z).sqr() - Z1Z1 - Z2Z2) * H;
        return r;
    }
    G1J neg() const { G1J r = *this; r.y = y.neg(); return r; }
    G1J mul(const Fr64& k) const {   // MSB-first double-and-add over the canonical scalar
        Fr64 c = k.from_mont();
        G1J acc = identity();
        for (int i = 255; i >= 0; --i) { acc = acc.dbl(); if ((c.l[i >> 6] >> (i & 63)) & 1) acc = acc.add(*this); }
        return acc;
    }
    bool to_affine(Fq64& ax, Fq64& ay) const {   // false for the identity (affine (0,0))
        if (is_identity()) { ax = Fq64::zero(); ay = Fq64::zero(); return false; }
        Fq64 zi = z.inverse(), zi2 = zi.sqr();
        ax = x * zi2; ay = y * zi2 * zi;
        return true;
    }
    // coordinates reduced (< p) and Y^2 = X^3 + 3 Z^6 (the identity, Z = 0, passes).  BN254 G1 has cofactor 1: on the curve means in the group.
    // arkworks' G1Projective cannot hold anything else; bytes that arrive over the C ABI can.
    bool is_valid() const {
        if (Fq64::geq_mod(x.l) || Fq64::geq_mod(y.l) || Fq64::geq_mod(z.l)) return false;
        if (is_identity()) return true;
        Fq64 z2 = z.sqr(), z6 = z2.sqr() * z2;
        return y.sqr() == x.sqr() * x + (z6.dbl() + z6);
    }
    bool equals(const G1J& o) const {
        if (is_identity() || o.is_identity()) return is_identity() && o.is_identity();
        Fq64 a = z.sqr(), b = o.z.sqr();
        return x * b == o.x * a && y * o.z * b == o.y * z * a;
    }
};

// KZGCommitmentValue::hash (src/commitments.rs:73-84): affine x, canonical LE bytes, reduced mod r
inline Fr64 g1_hash(const G1J& p) {
    Fq64 ax, ay;
    if (!p.to_affine(ax, ay)) return Fr64::zero();
    Fq64 c = ax.from_mont();
    uint64_t v[4]; memcpy(v, c.l, 32);
    while (Fr64::geq_mod(v)) Fr64::sub_mod(v);   // x < p < 2r: at most one subtraction
    Fr64 r; memcpy(r.l, v, 32);
    return r * Fr64::r2();
}

// ark-serialize 0.4.2 compressed G1: x LE | 0x80 (y > -y) | 0x40 (infinity) in the top bits of byte 31
inline void g1_compress(const G1J& p, uint8_t out[32]) {
    Fq64 ax, ay;
    if (!p.to_affine(ax, ay)) { memset(out, 0, 32); out[31] |= 0x40; return; }
    Fq64 cx = ax.from_mont(), cy = ay.from_mont(), cny = ay.neg().from_mont();
    memcpy(out, cx.l, 32);
    bool greater = false;
    for (int i = 3; i >= 0; --i) { if (cy.l[i] > cny.l[i]) { greater = true; break; } if (cy.l[i] < cny.l[i]) break; }
    if (greater) out[31] |= 0x80;
}

}  // namespace host
}  // namespace tsg
