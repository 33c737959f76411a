// host/pairing.hpp - BN254 optimal-ate pairing on the CPU, for KZGCommitment::verify (src/commitments.rs:201-228),
// KZGCommitment::batch_verify (:230-301) and through them Twist::verify / Shout::verify (src/twist.rs:255-304,
// src/shout.rs:225-274).  Verification is host-only work in the reference (two pairings per opening, tens of ms);
// it is not a GPU target.
//
// ark-bn254 0.4.0 parameters: Fq2 = Fq[u]/(u^2 + 1), xi = 9 + u, Fq12 = Fq2[w]/(w^6 - xi), D-type twist
// E': y^2 = x^3 + 3/xi, curve parameter x = 4965661367192848881, Miller loop over 6x + 2 followed by the two
// Frobenius line additions, final exponentiation by (p^12 - 1)/r split into its easy part (conjugation, inversion, Frobenius) and a
// square-and-multiply over (p^4 - p^2 + 1)/r.  Deliberately simple (affine twist arithmetic in the Miller loop, schoolbook Fq12): ~10 ms per
// pairing product, correctness first.
// Only the BOOLEAN result of a verification is observable, and it is the same for any correct bilinear pairing.
#pragma once
#include <vector>
#include "field64.hpp"

namespace tsg {
namespace host {

struct Fq2 {
    Fq64 c0, c1;   // c0 + c1 u, u^2 = -1
    static Fq2 zero() { return {Fq64::zero(), Fq64::zero()}; }
    static Fq2 one() { return {Fq64::one(), Fq64::zero()}; }
    static Fq2 from_fq(const Fq64& a) { return {a, Fq64::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
    Fq2 operator+(const Fq2& o) const { return {c0 + o.c0, c1 + o.c1}; }
    Fq2 operator-(const Fq2& o) const { return {c0 - o.c0, c1 - o.c1}; }
    Fq2 neg() const { return {c0.neg(), c1.neg()}; }
    Fq2 conj() const { return {c0, c1.neg()}; }
    Fq2 operator*(const Fq2& o) const {
        Fq64 a = c0 * o.c0, b = c1 * o.c1;
        Fq64 c = (c0 + c1) * (o.c0 + o.c1);
        return {a - b, c - a - b};
    }
    Fq2 mul_fq(const Fq64& s) const { return {c0 * s, c1 * s}; }
    Fq2 sqr() const { return *this * *this; }
    Fq2 dbl() const { return *this + *this; }
    Fq2 inverse() const {   // (c0 - c1 u) / (c0^2 + c1^2)
        Fq64 n = (c0.sqr() + c1.sqr()).inverse();
        return {c0 * n, (c1 * n).neg()};
    }
    Fq2 pow(const uint64_t* e, int nlimbs) const {
        Fq2 acc = one();
        for (int i = nlimbs * 64 - 1; i >= 0; --i) { acc = acc.sqr(); if ((e[i >> 6] >> (i & 63)) & 1) acc = acc * *this; }
        return acc;
    }
};

inline Fq2 xi() { return {Fq64::from_u64(9), Fq64::from_u64(1)}; }

// Fq12 as a degree-6 extension of Fq2: sum c[i] w^i, w^6 = xi
struct Fq12 {
    Fq2 c[6];
    static Fq12 one() { Fq12 r; for (auto& x : r.c) x = Fq2::zero(); r.c[0] = Fq2::one(); return r; }
    bool operator==(const Fq12& o) const { for (int i = 0; i < 6; ++i) if (!(c[i] == o.c[i])) return false; return true; }
    bool is_one() const { return *this == one(); }
    Fq12 operator*(const Fq12& o) const {
        Fq2 t[11];
        for (auto& x : t) x = Fq2::zero();
        for (int i = 0; i < 6; ++i) {
            if (c[i].is_zero()) continue;
            for (int j = 0; j < 6; ++j) { if (o.c[j].is_zero()) continue; t[i + j] = t[i + j] + c[i] * o.c[j]; }
        }
        Fq12 r; const Fq2 x = xi();
        for (int i = 0; i < 6; ++i) r.c[i] = t[i];
        for (int i = 6; i < 11; ++i) r.c[i - 6] = r.c[i - 6] + t[i] * x;
        return r;
    }
    Fq12 sqr() const {   // symmetric schoolbook: 21 instead of 36 Fq2 products
        Fq2 t[11];
        for (auto& x : t) x = Fq2::zero();
        for (int i = 0; i < 6; ++i) {
            if (c[i].is_zero()) continue;
            t[2 * i] = t[2 * i] + c[i].sqr();
            for (int j = i + 1; j < 6; ++j) { if (c[j].is_zero()) continue; t[i + j] = t[i + j] + (c[i] * c[j]).dbl(); }
        }
        Fq12 r; const Fq2 x = xi();
        for (int i = 0; i < 6; ++i) r.c[i] = t[i];
        for (int i = 6; i < 11; ++i) r.c[i - 6] = r.c[i - 6] + t[i] * x;
        return r;
    }
    // f^(p^6): w^(p^6) = -w and Fq2 is fixed, so the odd coefficients change sign
    Fq12 conj6() const { Fq12 r = *this; r.c[1] = c[1].neg(); r.c[3] = c[3].neg(); r.c[5] = c[5].neg(); return r; }
    // f^p: coefficient i -> conj(c_i) * gamma^i, gamma = xi^((p - 1) / 6)
    Fq12 frobenius() const {
        static const Fq2 gamma = [] { uint64_t e[4]; for (int i = 0; i < 4; ++i) e[i] = Fq64::modl(i); e[0] -= 1;
                                      unsigned __int128 rem = 0; for (int i = 3; i >= 0; --i) { unsigned __int128 cur = (rem << 64) | e[i]; e[i] = (uint64_t)(cur / 6); rem = cur % 6; }
                                      return xi().pow(e, 4); }();
        Fq12 r; Fq2 g = Fq2::one();
        for (int i = 0; i < 6; ++i) { r.c[i] = c[i].conj() * g; g = g * gamma; }
        return r;
    }
    // inverse through Fq12 = Fq6[w] / (w^2 - v), Fq6 = Fq2[v] / (v^3 - xi), v = w^2:  f = a + w b,  1/f = (a - w b) / (a^2 - v b^2)
    Fq12 inverse() const {
        struct F6 {
            Fq2 a0, a1, a2;
            F6 mul(const F6& o) const {
                const Fq2 x = xi();
                return {a0 * o.a0 + (a1 * o.a2 + a2 * o.a1) * x, a0 * o.a1 + a1 * o.a0 + a2 * o.a2 * x, a0 * o.a2 + a1 * o.a1 + a2 * o.a0};
            }
            F6 sub(const F6& o) const { return {a0 - o.a0, a1 - o.a1, a2 - o.a2}; }
            F6 mul_v() const { return {a2 * xi(), a0, a1}; }
            F6 inverse() const {
                const Fq2 x = xi();
                Fq2 t0 = a0.sqr() - a1 * a2 * x, t1 = a2.sqr() * x - a0 * a1, t2 = a1.sqr() - a0 * a2;
                Fq2 d = (a0 * t0 + (a2 * t1 + a1 * t2) * x).inverse();
                return {t0 * d, t1 * d, t2 * d};
            }
        };
        F6 a = {c[0], c[2], c[4]}, b = {c[1], c[3], c[5]};
        F6 n = a.mul(a).sub(b.mul(b).mul_v()).inverse();
        F6 ra = a.mul(n), rb = b.mul(n);
        Fq12 r;
        r.c[0] = ra.a0; r.c[2] = ra.a1; r.c[4] = ra.a2;
        r.c[1] = rb.a0.neg(); r.c[3] = rb.a1.neg(); r.c[5] = rb.a2.neg();
        return r;
    }
    Fq12 pow(const uint64_t* e, int nlimbs) const {
        Fq12 acc = one();
        bool started = false;
        for (int i = nlimbs * 64 - 1; i >= 0; --i) {
            if (started) acc = acc.sqr();
            if ((e[i >> 6] >> (i & 63)) & 1) { acc = started ? acc * *this : *this; started = true; }
        }
        return acc;
    }
};

// affine point on the twist E'(Fq2): y^2 = x^3 + 3/xi
struct G2A {
    Fq2 x, y;
    bool inf;
    static G2A infinity() { return {Fq2::zero(), Fq2::zero(), true}; }
    static G2A generator();
    G2A neg() const { return {x, y.neg(), inf}; }
    bool on_curve() const {
        if (inf) return true;
        Fq2 b = Fq2::from_fq(Fq64::from_u64(3)) * xi().inverse();
        return y.sqr() == x.sqr() * x + b;
    }
    // returns the sum and the slope used (lambda); `vertical` is set when the result is infinity
    G2A add(const G2A& o) const {
        if (inf) return o;
        if (o.inf) return *this;
        Fq2 lam;
        if (x == o.x) {
            if (!(y == o.y) || y.is_zero()) return infinity();
            lam = (x.sqr().dbl() + x.sqr()) * y.dbl().inverse();
        } else {
            lam = (o.y - y) * (o.x - x).inverse();
        }
        Fq2 x3 = lam.sqr() - x - o.x;
        return {x3, lam * (x - x3) - y, false};
    }
    // scalar multiplication in Jacobian coordinates (a = 0: dbl-2009-l, madd-2007-bl), one inversion at the end instead of one per step
    G2A mul(const Fr64& k) const {
        if (inf) return *this;
        Fr64 c = k.from_mont();
        Fq2 X = Fq2::zero(), Y = Fq2::one(), Z = Fq2::zero();   // identity: Z = 0
        for (int i = 255; i >= 0; --i) {
            if (!Z.is_zero()) {                                   // double
                Fq2 A = X.sqr(), B = Y.sqr(), Cc = B.sqr();
                Fq2 D = ((X + B).sqr() - A - Cc).dbl();
                Fq2 E = A.dbl() + A, F = E.sqr();
                Fq2 X3 = F - D.dbl();
                Fq2 Y3 = E * (D - X3) - Cc.dbl().dbl().dbl();
                Fq2 Z3 = (Y * Z).dbl();
                X = X3; Y = Y3; Z = Z3;
            }
            if ((c.l[i >> 6] >> (i & 63)) & 1) {
                if (Z.is_zero()) { X = x; Y = y; Z = Fq2::one(); continue; }
                Fq2 Z1Z1 = Z.sqr(), U2 = x * Z1Z1, S2 = y * Z * Z1Z1;
                if (U2 == X) {
                    if (S2 == Y) {                                // acc == this: double (same formulas)
                        Fq2 A = X.sqr(), B = Y.sqr(), Cc = B.sqr();
                        Fq2 D = ((X + B).sqr() - A - Cc).dbl();
                        Fq2 E = A.dbl() + A, F = E.sqr();
                        Fq2 X3 = F - D.dbl();
                        Fq2 Y3 = E * (D - X3) - Cc.dbl().dbl().dbl();
                        Fq2 Z3 = (Y * Z).dbl();
                        X = X3; Y = Y3; Z = Z3;
                    } else { X = Fq2::zero(); Y = Fq2::one(); Z = Fq2::zero(); }
                    continue;
                }
                Fq2 H = U2 - X, HH = H.sqr(), I = HH.dbl().dbl(), J = H * I, rr = (S2 - Y).dbl(), V = X * I;
                Fq2 X3 = rr.sqr() - J - V.dbl();
                Fq2 Y3 = rr * (V - X3) - (Y * J).dbl();
                Fq2 Z3 = (Z + H).sqr() - Z1Z1 - HH;
                X = X3; Y = Y3; Z = Z3;
            }
        }
        if (Z.is_zero()) return infinity();
        Fq2 zi = Z.inverse(), zi2 = zi.sqr();
        return {X * zi2, Y * zi2 * zi, false};
    }
};

inline Fq64 fq_from_decimal(const char* s) {
    Fq64 acc = Fq64::zero(), ten = Fq64::from_u64(10);
    for (; *s; ++s) acc = acc * ten + Fq64::from_u64((uint64_t)(*s - '0'));
    return acc;
}
inline G2A G2A::generator() {   // ark-bn254 G2 generator
    G2A g;
    g.x = {fq_from_decimal("10857046999023057135944570762232829481370756359578518086990519993285655852781"),
           fq_from_decimal("11559732032986387107991004021392285783925812861821192530917403151452391805634")};
    g.y = {fq_from_decimal("8495653923123431417604973247489272438418190587263600148770280649306958101930"),
           fq_from_decimal("4082367875863433681332203403145435568316851327593401208105741076214120093531")};
    g.inf = false;
    return g;
}

// line through T and Q (tangent when T == Q) on the twist, evaluated at P = (xP, yP) in E(Fq):
//   l = yP - lambda xP w + (lambda x_T - y_T) w^3        (untwist: (x, y) -> (x w^2, y w^3))
// advances T to T + Q.  Returns false (and leaves f untouched) for a vertical line, which cannot occur for points
// of prime order r inside the loop.
inline bool line_and_add(G2A& T, const G2A& Q, const Fq64& xP, const Fq64& yP, Fq12& l) {
    Fq2 lam;
    if (T.x == Q.x) {
        if (!(T.y == Q.y) || T.y.is_zero()) { T = G2A::infinity(); return false; }
        Fq2 xx = T.x.sqr();
        lam = (xx.dbl() + xx) * T.y.dbl().inverse();
    } else {
        lam = (Q.y - T.y) * (Q.x - T.x).inverse();
    }
    for (auto& x : l.c) x = Fq2::zero();
    l.c[0] = Fq2::from_fq(yP);
    l.c[1] = lam.mul_fq(xP).neg();
    l.c[3] = lam * T.x - T.y;
    Fq2 x3 = lam.sqr() - T.x - Q.x;
    Fq2 y3 = lam * (T.x - x3) - T.y;
    T = {x3, y3, false};
    return true;
}

// Miller loop f_{6x+2,Q}(P) with the two Frobenius corrections; P affine in G1 (Montgomery Fq), Q affine on the twist
inline Fq12 miller_loop(const Fq64& xP, const Fq64& yP, bool p_inf, const G2A& Q) {
    Fq12 f = Fq12::one();
    if (p_inf || Q.inf) return f;
    // 6x + 2 = 29793968203157093288 = 0x1_9d797039be763ba8 (65 bits)
    const uint64_t lo = 0x9d797039be763ba8ull;
    G2A T = Q;
    Fq12 l;
    for (int i = 63; i >= 0; --i) {   // bit 64 is the leading one
        f = f.sqr();
        G2A T2 = T;
        if (line_and_add(T, T2, xP, yP, l)) f = f * l;
        if ((lo >> i) & 1) { if (line_and_add(T, Q, xP, yP, l)) f = f * l; }
    }
    // Q1 = pi(Q), Q2 = pi^2(Q): x -> conj^k(x) xi^((p^k-1)/3), y -> conj^k(y) xi^((p^k-1)/2)
    static const Fq2 g12 = [] { uint64_t e[4]; for (int i = 0; i < 4; ++i) e[i] = Fq64::modl(i); e[0] -= 1;
                                // (p - 1) / 3
                                unsigned __int128 rem = 0; for (int i = 3; i >= 0; --i) { unsigned __int128 cur = (rem << 64) | e[i]; e[i] = (uint64_t)(cur / 3); rem = cur % 3; }
                                return xi().pow(e, 4); }();
    static const Fq2 g13 = [] { uint64_t e[4]; for (int i = 0; i < 4; ++i) e[i] = Fq64::modl(i); e[0] -= 1;
                                for (int i = 0; i < 4; ++i) e[i] = (e[i] >> 1) | (i < 3 ? e[i + 1] << 63 : 0);   // (p - 1) / 2
                                return xi().pow(e, 4); }();
    G2A Q1 = {Q.x.conj() * g12, Q.y.conj() * g13, false};
    // pi^2: apply pi twice
    G2A Q2 = {Q1.x.conj() * g12, Q1.y.conj() * g13, false};
    if (line_and_add(T, Q1, xP, yP, l)) f = f * l;
    G2A nQ2 = Q2.neg();
    if (line_and_add(T, nQ2, xP, yP, l)) f = f * l;
    return f;
}

// f^((p^12 - 1) / r) = ((f^(p^6 - 1))^(p^2 + 1))^((p^4 - p^2 + 1) / r): the easy part by one conjugation, one inversion and two Frobenius maps,
// the hard part by square-and-multiply over its 761-bit exponent (exactly the same GT element as a plain power by (p^12 - 1) / r)
inline Fq12 final_exponentiation(const Fq12& f) {
    static const uint64_t E[12] = {
        0xe81bb482ccdf42b1ull, 0x5abf5cc4f49c36d4ull, 0xf1154e7e1da014fdull, 0xdcc7b44c87cdbacfull, 0xaaa441e3954bcf8aull, 0x6b887d56d5095f23ull,
        0x79581e16f3fd90c6ull, 0x3b1b1355d189227dull, 0x4e529a5861876f6bull, 0x6c0eb522d5b12278ull, 0x331ec15183177fafull, 0x01baaa710b0759adull};   // (p^4 - p^2 + 1) / r
    Fq12 t = f.conj6() * f.inverse();          // f^(p^6 - 1)
    t = t.frobenius().frobenius() * t;         // ^(p^2 + 1)
    return t.pow(E, 12);
}
// the plain power by (p^12 - 1) / r (44 limbs): kept as the cross-check of the split form (tests/test_host_pairing.py)
inline Fq12 final_exponentiation_plain(const Fq12& f) {
    static const uint64_t E[44] = {
        0x86964b64ca86f120ull, 0x40a4efb7e54523a4ull, 0x837fa97896e84abbull, 0x361102b6b9b2b918ull, 0xc0de81def35692daull, 0xbe04c7e8a6c3c760ull,
        0xd766f9c9d570bb7full, 0xc230974d83561841ull, 0x5bba1668c3be69a3ull, 0x7f3811c410526294ull, 0x29baee7ddadda71cull, 0xbf813b8d145da900ull,
        0x641bbadf423f9a2cull, 0xa80bb4ea44eacc5eull, 0xcd65664814fde37cull, 0x4a0364b9580291d2ull, 0xee93dfb10826f0ddull, 0x6b42db8dc5514724ull,
        0xbb10cf430b0f3785ull, 0x40494e406f804216ull, 0x55cfe107acf3aafbull, 0x2088ec80e0ebae87ull, 0x846a3ed011a337a0ull, 0x48a45a4a1e3a5195ull,
        0xe5664568dfc50e16ull, 0xab6a41294c0cc4ebull, 0x82d0d602d268c7daull, 0x6668449aed3cc48aull, 0x5062cd0fb2015dfcull, 0x7f2940a8b1ddb3d1ull,
        0x77f5b63a2a226448ull, 0xfef0781361e443aeull, 0xf977870e88d5c6c8ull, 0x790364a61f676baaull, 0x5887e72eceaddea3ull, 0x1377e563a09a1b70ull,
        0x0c54efee1bd8c3b2ull, 0x3ec3d15ad524d8f7ull, 0xdaf15466b2383a5dull, 0xe1e30a73bb94fec0ull, 0x6a1c71015f3f7be2ull, 0x842d43bf6369b1ffull,
        0x20fddadf107d20bcull, 0x0000002f4b6dc970ull};   // (p^12 - 1) / r
    return f.pow(E, 44);
}

// e(P, Q) in GT (P Jacobian G1, Q affine twist point)
inline Fq12 pairing(const G1J& P, const G2A& Q) {
    Fq64 ax, ay;
    bool ok = P.to_affine(ax, ay);
    return final_exponentiation(miller_loop(ax, ay, !ok, Q));
}
// prod_i e(P_i, Q_i) == 1 with a single final exponentiation
inline bool pairing_product_is_one(const std::vector<G1J>& Ps, const std::vector<G2A>& Qs) {
    Fq12 f = Fq12::one();
    for (size_t i = 0; i < Ps.size(); ++i) {
        Fq64 ax, ay;
        bool ok = Ps[i].to_affine(ax, ay);
        f = f * miller_loop(ax, ay, !ok, Qs[i]);
    }
    return final_exponentiation(f).is_one();
}

// CommitmentVerificationKey (src/utils.rs:66-76)
struct VerifyKey {
    G1J g1_generator;
    G2A g2_generator, g2_tau;
};

// KZGCommitment::verify (src/commitments.rs:201-228): e(C - [v]_1, [1]_2) == e(pi, [tau]_2 - [z]_2)
inline bool kzg_verify(const VerifyKey& vk, const G1J& commitment, const Fr64& point, const Fr64& value, const G1J& proof) {
    G1J left = commitment.add(vk.g1_generator.mul(value).neg());
    G2A right = vk.g2_tau.add(vk.g2_generator.mul(point).neg());
    // e(left, g2) == e(proof, right)  <=>  e(left, g2) * e(-proof, right) == 1
    return pairing_product_is_one({left, proof.neg()}, {vk.g2_generator, right});
}

}  // namespace host
}  // namespace tsg
