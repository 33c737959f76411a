// fp.cuh - 254-bit prime-field arithmetic (BN254 Fr and Fq) on 8 x 32-bit limbs, Montgomery form.
//
// Memory layout of an element is the reference's: ark_bn254::Fr / Fq = BigInt<4> = uint64_t[4]
// little-endian limbs holding a*2^256 mod p (reference src/utils.rs:14) - i.e. exactly
// these 8 little-endian uint32_t limbs.  No conversion happens at the boundary.
//
// Multiplication is an interleaved (CIOS-style) Montgomery product on two half-width accumulators:
// products of even-indexed limbs of the multiplicand land on one carry chain, odd-indexed ones on a
// second chain one limb higher, so every 32x32->64 multiply-accumulate is a carry-chained
// IMAD.WIDE (mad.lo.cc / madc.hi.cc pair) and no partial product ever needs a ripple carry.
// After each word step the low accumulator limb is zero and the two accumulators swap roles.
// Cost: 128 wide MACs + 8 low multiplies + ~40 adds per product.
#pragma once
#include "ptx.cuh"

namespace tsg {

struct FrP {   // BN254 scalar field r
    static constexpr uint32_t INV = 0xefffffffu;   // -r^-1 mod 2^32
    TSG_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    TSG_HD static constexpr uint32_t one(int i) {   // 2^256 mod r
        constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    TSG_HD static constexpr uint32_t r2(int i) {    // 2^512 mod r
        constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return m[i];
    }
};
struct FqP {   // BN254 base field p
    static constexpr uint32_t INV = 0xe4866389u;   // -p^-1 mod 2^32
    TSG_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    TSG_HD static constexpr uint32_t one(int i) {
        constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    TSG_HD static constexpr uint32_t r2(int i) {
        constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return m[i];
    }
};

namespace limb {

// r = a - mod if a >= mod else a   (a < 2*mod)
template <class P>
TSG_HD void cond_sub_mod(uint32_t* a) {
    uint32_t t[8];
    t[0] = ptx::sub_cc(a[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < 8; ++i) t[i] = ptx::subc_cc(a[i], P::mod(i));
    uint32_t borrow = ptx::subc(0u, 0u);   // 0 or 0xffffffff
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = borrow ? a[i] : t[i];
}

template <class P>
TSG_HD void add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    r[0] = ptx::add_cc(a[0], b[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) r[i] = ptx::addc_cc(a[i], b[i]);
    r[7] = ptx::addc(a[7], b[7]);   // a + b < 2^255: no carry out
    cond_sub_mod<P>(r);
}

template <class P>
TSG_HD void sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t t[8];
    t[0] = ptx::sub_cc(a[0], b[0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) t[i] = ptx::subc_cc(a[i], b[i]);
    uint32_t borrow = ptx::subc(0u, 0u);
    // add back mod & borrow
    r[0] = ptx::add_cc(t[0], P::mod(0) & borrow);
#pragma unroll
    for (int i = 1; i < 7; ++i) r[i] = ptx::addc_cc(t[i], P::mod(i) & borrow);
    r[7] = ptx::addc(t[7], P::mod(7) & borrow);
}

// ---- one word step of the interleaved Montgomery product ------------------------------------------
// Y: accumulator aligned at limb 0, X: accumulator aligned at limb 1 (value = Y + X * 2^32).
// Adds m * mod with m chosen so that the low limb cancels.
template <class P>
TSG_HD void redc_step(uint32_t* X, uint32_t* Y) {
    uint32_t m = ptx::mul_lo(Y[0], P::INV);
    X[0] = ptx::mad_lo_cc(P::mod(1), m, X[0]);
    X[1] = ptx::madc_hi_cc(P::mod(1), m, X[1]);
    X[2] = ptx::madc_lo_cc(P::mod(3), m, X[2]);
    X[3] = ptx::madc_hi_cc(P::mod(3), m, X[3]);
    X[4] = ptx::madc_lo_cc(P::mod(5), m, X[4]);
    X[5] = ptx::madc_hi_cc(P::mod(5), m, X[5]);
    X[6] = ptx::madc_lo_cc(P::mod(7), m, X[6]);
    X[7] = ptx::madc_hi(P::mod(7), m, X[7]);
    Y[0] = ptx::mad_lo_cc(P::mod(0), m, Y[0]);
    Y[1] = ptx::madc_hi_cc(P::mod(0), m, Y[1]);
    Y[2] = ptx::madc_lo_cc(P::mod(2), m, Y[2]);
    Y[3] = ptx::madc_hi_cc(P::mod(2), m, Y[3]);
    Y[4] = ptx::madc_lo_cc(P::mod(4), m, Y[4]);
    Y[5] = ptx::madc_hi_cc(P::mod(4), m, Y[5]);
    Y[6] = ptx::madc_lo_cc(P::mod(6), m, Y[6]);
    Y[7] = ptx::madc_hi_cc(P::mod(6), m, Y[7]);
    X[7] = ptx::addc(X[7], 0u);
}

// r = a * b * 2^-256 mod p, inputs < p, output < p
template <class P>
TSG_HD void mont_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t acc[2][8];
    {
        uint32_t* Y = acc[0]; uint32_t* X = acc[1];
        uint32_t bi = b[0];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            ptx::mul_wide(Y[j], Y[j + 1], a[j], bi);
            ptx::mul_wide(X[j], X[j + 1], a[j + 1], bi);
        }
        redc_step<P>(X, Y);
    }
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        uint32_t* Y = acc[i & 1]; uint32_t* X = acc[(i & 1) ^ 1];
        uint32_t bi = b[i];
        // shift: previous low accumulator (now X, with X[0] == 0) moves down two limbs; X[1] joins Y[0]
        Y[0] = ptx::add_cc(Y[0], X[1]);
        X[0] = ptx::madc_lo_cc(a[1], bi, X[2]);
        X[1] = ptx::madc_hi_cc(a[1], bi, X[3]);
        X[2] = ptx::madc_lo_cc(a[3], bi, X[4]);
        X[3] = ptx::madc_hi_cc(a[3], bi, X[5]);
        X[4] = ptx::madc_lo_cc(a[5], bi, X[6]);
        X[5] = ptx::madc_hi_cc(a[5], bi, X[7]);
        X[6] = ptx::madc_lo_cc(a[7], bi, 0u);
        X[7] = ptx::madc_hi(a[7], bi, 0u);
        Y[0] = ptx::mad_lo_cc(a[0], bi, Y[0]);
        Y[1] = ptx::madc_hi_cc(a[0], bi, Y[1]);
        Y[2] = ptx::madc_lo_cc(a[2], bi, Y[2]);
        Y[3] = ptx::madc_hi_cc(a[2], bi, Y[3]);
        Y[4] = ptx::madc_lo_cc(a[4], bi, Y[4]);
        Y[5] = ptx::madc_hi_cc(a[4], bi, Y[5]);
        Y[6] = ptx::madc_lo_cc(a[6], bi, Y[6]);
        Y[7] = ptx::madc_hi_cc(a[6], bi, Y[7]);
        X[7] = ptx::addc(X[7], 0u);
        redc_step<P>(X, Y);
    }
    // after step 7: Y = acc[1] (Y[0] == 0), X = acc[0]; result = X + (Y >> 32)
    {
        uint32_t* Y = acc[1]; uint32_t* X = acc[0];
        r[0] = ptx::add_cc(X[0], Y[1]);
#pragma unroll
        for (int k = 1; k < 7; ++k) r[k] = ptx::addc_cc(X[k], Y[k + 1]);
        r[7] = ptx::addc(X[7], 0u);
    }
    cond_sub_mod<P>(r);
}

// t[0..16) = a * b (full 512-bit product, no reduction)
TSG_HD void mul_wide(uint32_t* t, const uint32_t* a, const uint32_t* b) {
    uint32_t E[16], O[16];
#pragma unroll
    for (int k = 8; k < 16; ++k) { E[k] = 0; O[k] = 0; }
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        ptx::mul_wide(E[j], E[j + 1], a[j], b[0]);
        ptx::mul_wide(O[j], O[j + 1], a[j + 1], b[0]);
    }
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        uint32_t bi = b[i];
        if (i & 1) {
            // even a-limbs land on odd product limbs: O index (i + j - 1)
            O[i - 1] = ptx::mad_lo_cc(a[0], bi, O[i - 1]);
            O[i + 0] = ptx::madc_hi_cc(a[0], bi, O[i + 0]);
            O[i + 1] = ptx::madc_lo_cc(a[2], bi, O[i + 1]);
            O[i + 2] = ptx::madc_hi_cc(a[2], bi, O[i + 2]);
            O[i + 3] = ptx::madc_lo_cc(a[4], bi, O[i + 3]);
            O[i + 4] = ptx::madc_hi_cc(a[4], bi, O[i + 4]);
            O[i + 5] = ptx::madc_lo_cc(a[6], bi, O[i + 5]);
            O[i + 6] = ptx::madc_hi_cc(a[6], bi, O[i + 6]);
            if (i + 7 < 15) O[i + 7] = ptx::addc(O[i + 7], 0u);
            // odd a-limbs land on even product limbs: E index (i + j)
            E[i + 1] = ptx::mad_lo_cc(a[1], bi, E[i + 1]);
            E[i + 2] = ptx::madc_hi_cc(a[1], bi, E[i + 2]);
            E[i + 3] = ptx::madc_lo_cc(a[3], bi, E[i + 3]);
            E[i + 4] = ptx::madc_hi_cc(a[3], bi, E[i + 4]);
            E[i + 5] = ptx::madc_lo_cc(a[5], bi, E[i + 5]);
            E[i + 6] = ptx::madc_hi_cc(a[5], bi, E[i + 6]);
            E[i + 7] = ptx::madc_lo_cc(a[7], bi, E[i + 7]);
            E[i + 8] = ptx::madc_hi_cc(a[7], bi, E[i + 8]);
            if (i + 9 < 16) E[i + 9] = ptx::addc(E[i + 9], 0u);
        } else {
            E[i + 0] = ptx::mad_lo_cc(a[0], bi, E[i + 0]);
            E[i + 1] = ptx::madc_hi_cc(a[0], bi, E[i + 1]);
            E[i + 2] = ptx::madc_lo_cc(a[2], bi, E[i + 2]);
            E[i + 3] = ptx::madc_hi_cc(a[2], bi, E[i + 3]);
            E[i + 4] = ptx::madc_lo_cc(a[4], bi, E[i + 4]);
            E[i + 5] = ptx::madc_hi_cc(a[4], bi, E[i + 5]);
            E[i + 6] = ptx::madc_lo_cc(a[6], bi, E[i + 6]);
            E[i + 7] = ptx::madc_hi_cc(a[6], bi, E[i + 7]);
            if (i + 8 < 16) E[i + 8] = ptx::addc(E[i + 8], 0u);
            O[i + 0] = ptx::mad_lo_cc(a[1], bi, O[i + 0]);
            O[i + 1] = ptx::madc_hi_cc(a[1], bi, O[i + 1]);
            O[i + 2] = ptx::madc_lo_cc(a[3], bi, O[i + 2]);
            O[i + 3] = ptx::madc_hi_cc(a[3], bi, O[i + 3]);
            O[i + 4] = ptx::madc_lo_cc(a[5], bi, O[i + 4]);
            O[i + 5] = ptx::madc_hi_cc(a[5], bi, O[i + 5]);
            O[i + 6] = ptx::madc_lo_cc(a[7], bi, O[i + 6]);
            O[i + 7] = ptx::madc_hi_cc(a[7], bi, O[i + 7]);
            if (i + 8 < 15) O[i + 8] = ptx::addc(O[i + 8], 0u);
        }
    }
    t[0] = E[0];
    t[1] = ptx::add_cc(E[1], O[0]);
#pragma unroll
    for (int k = 2; k < 15; ++k) t[k] = ptx::addc_cc(E[k], O[k - 1]);
    t[15] = ptx::addc(E[15], O[14]);
}

// acc[0..16) += t[0..16)   (caller guarantees no overflow past 512 bits)
TSG_HD void wide_add(uint32_t* acc, const uint32_t* t) {
    acc[0] = ptx::add_cc(acc[0], t[0]);
#pragma unroll
    for (int k = 1; k < 15; ++k) acc[k] = ptx::addc_cc(acc[k], t[k]);
    acc[15] = ptx::addc(acc[15], t[15]);
}

// Bring the high half of a 512-bit accumulator below p (value stays congruent mod p * 2^256 ... i.e. the
// represented residue mod p is unchanged because multiples of p*2^256 are removed).  Precondition:
// acc < 2^512.  Postcondition: acc < p * 2^256, so 16 more products of reduced operands fit again.
template <class P>
TSG_HD void wide_normalize(uint32_t* acc) {
    // high < 2^256 < 6p: subtract 4p, 2p, p conditionally
#pragma unroll
    for (int sh = 2; sh >= 0; --sh) {
        uint32_t t[8];
        // (p << sh) limbs
        uint32_t ps[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ps[i] = (P::mod(i) << sh) | (sh && i ? (P::mod(i - 1) >> (32 - sh)) : 0u);
        t[0] = ptx::sub_cc(acc[8], ps[0]);
#pragma unroll
        for (int i = 1; i < 8; ++i) t[i] = ptx::subc_cc(acc[8 + i], ps[i]);
        uint32_t borrow = ptx::subc(0u, 0u);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[8 + i] = borrow ? acc[8 + i] : t[i];
    }
}

// r = T * 2^-256 mod p for T[0..16) < p * 2^256; output < p
template <class P>
TSG_HD void mont_reduce(uint32_t* r, const uint32_t* T) {
    uint32_t acc[2][8];
    {
        uint32_t* Y = acc[0]; uint32_t* X = acc[1];
#pragma unroll
        for (int k = 0; k < 8; ++k) { Y[k] = T[k]; X[k] = 0; }
        redc_step<P>(X, Y);
    }
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        uint32_t* Y = acc[i & 1]; uint32_t* X = acc[(i & 1) ^ 1];
        Y[0] = ptx::add_cc(Y[0], X[1]);
#pragma unroll
        for (int k = 0; k < 6; ++k) X[k] = ptx::addc_cc(X[k + 2], 0u);
        X[6] = ptx::addc_cc(T[7 + i], 0u);
        X[7] = ptx::addc(0u, 0u);
        redc_step<P>(X, Y);
    }
    {
        uint32_t* Y = acc[1]; uint32_t* X = acc[0];
        r[0] = ptx::add_cc(X[0], Y[1]);
#pragma unroll
        for (int k = 1; k < 7; ++k) r[k] = ptx::addc_cc(X[k], Y[k + 1]);
        r[7] = ptx::addc(X[7], T[15]);
    }
    cond_sub_mod<P>(r);
}


// r = (a b - c d) * 2^-256 mod p with ONE Montgomery reduction: the two 512-bit products are subtracted first (p * 2^256 is added back when the
// difference is negative, which leaves the residue unchanged), so 2 x 64 + 72 multiply-adds replace the 2 x 136 of two Montgomery products.
// Inputs < p, output < p: both products are < p^2 < p * 2^256, hence the adjusted difference is in [0, p * 2^256) as mont_reduce requires.
template <class P>
TSG_HD void mont_mul_sub(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d) {
    uint32_t W[16], V[16];
    mul_wide(W, a, b);
    mul_wide(V, c, d);
    W[0] = ptx::sub_cc(W[0], V[0]);
#pragma unroll
    for (int i = 1; i < 16; ++i) W[i] = ptx::subc_cc(W[i], V[i]);
    const uint32_t borrow = ptx::subc(0u, 0u);   // 0 or 0xffffffff
    W[8] = ptx::add_cc(W[8], P::mod(0) & borrow);
#pragma unroll
    for (int i = 1; i < 7; ++i) W[8 + i] = ptx::addc_cc(W[8 + i], P::mod(i) & borrow);
    W[15] = ptx::addc(W[15], P::mod(7) & borrow);
    mont_reduce<P>(r, W);
}

// ---- multiplication by a per-launch constant -------------------------------------------------------
// The fold of a sum-check round multiplies every table difference by the SAME challenge r.  With the
// eight residues  T[k] = r_canonical * 2^(32 k + 64) mod p  prepared once on the host,
//     a * r  ==  ( sum_k a_k T[k] ) * 2^-64   (mod p)          (a_k = 32-bit limbs of a)
// so the product needs the 64 limb products (all landing on the same ten limbs: no shifted rows), two
// Montgomery word steps (2 x 8 products) to strip the 2^64, and one conditional subtraction - about
// 80 wide multiply-adds instead of the 128 + 8 of a full Montgomery product.  For a in Montgomery
// form the result is the Montgomery form of the product: the same bits mont_mul(a, r_mont) returns.
// Bounds: sum < 8 * 2^32 * p; after the first word step < 9 p; after the second < 2 p.
template <class P>
TSG_HD void mul_ctab(uint32_t* r, const uint32_t* a, const uint32_t (*T)[8]) {
    uint32_t E[9], O[9];   // E aligned at limb 0, O at limb 1 (value = E + O * 2^32)
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        ptx::mul_wide(E[j], E[j + 1], a[0], T[0][j]);
        ptx::mul_wide(O[j], O[j + 1], a[0], T[0][j + 1]);
    }
    E[8] = 0; O[8] = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        const uint32_t ak = a[k];
        E[0] = ptx::mad_lo_cc(ak, T[k][0], E[0]);
        E[1] = ptx::madc_hi_cc(ak, T[k][0], E[1]);
        E[2] = ptx::madc_lo_cc(ak, T[k][2], E[2]);
        E[3] = ptx::madc_hi_cc(ak, T[k][2], E[3]);
        E[4] = ptx::madc_lo_cc(ak, T[k][4], E[4]);
        E[5] = ptx::madc_hi_cc(ak, T[k][4], E[5]);
        E[6] = ptx::madc_lo_cc(ak, T[k][6], E[6]);
        E[7] = ptx::madc_hi_cc(ak, T[k][6], E[7]);
        E[8] = ptx::addc(E[8], 0u);
        O[0] = ptx::mad_lo_cc(ak, T[k][1], O[0]);
        O[1] = ptx::madc_hi_cc(ak, T[k][1], O[1]);
        O[2] = ptx::madc_lo_cc(ak, T[k][3], O[2]);
        O[3] = ptx::madc_hi_cc(ak, T[k][3], O[3]);
        O[4] = ptx::madc_lo_cc(ak, T[k][5], O[4]);
        O[5] = ptx::madc_hi_cc(ak, T[k][5], O[5]);
        O[6] = ptx::madc_lo_cc(ak, T[k][7], O[6]);
        O[7] = ptx::madc_hi_cc(ak, T[k][7], O[7]);
        O[8] = ptx::addc(O[8], 0u);
    }
    uint32_t S[10];
    S[0] = E[0];
    S[1] = ptx::add_cc(E[1], O[0]);
#pragma unroll
    for (int k = 2; k < 9; ++k) S[k] = ptx::addc_cc(E[k], O[k - 1]);
    S[9] = ptx::addc(O[8], 0u);
    // two Montgomery word steps: S <- (S + m p) / 2^32 with m cancelling the low limb
#pragma unroll
    for (int st = 0; st < 2; ++st) {
        uint32_t* V = S + st;
        const uint32_t m = ptx::mul_lo(V[0], P::INV);
        V[0] = ptx::mad_lo_cc(P::mod(0), m, V[0]);
        V[1] = ptx::madc_hi_cc(P::mod(0), m, V[1]);
        V[2] = ptx::madc_lo_cc(P::mod(2), m, V[2]);
        V[3] = ptx::madc_hi_cc(P::mod(2), m, V[3]);
        V[4] = ptx::madc_lo_cc(P::mod(4), m, V[4]);
        V[5] = ptx::madc_hi_cc(P::mod(4), m, V[5]);
        V[6] = ptx::madc_lo_cc(P::mod(6), m, V[6]);
        V[7] = ptx::madc_hi_cc(P::mod(6), m, V[7]);
        if (st == 0) { V[8] = ptx::addc_cc(V[8], 0u); V[9] = ptx::addc(V[9], 0u); }
        else V[8] = ptx::addc(V[8], 0u);
        V[1] = ptx::mad_lo_cc(P::mod(1), m, V[1]);
        V[2] = ptx::madc_hi_cc(P::mod(1), m, V[2]);
        V[3] = ptx::madc_lo_cc(P::mod(3), m, V[3]);
        V[4] = ptx::madc_hi_cc(P::mod(3), m, V[4]);
        V[5] = ptx::madc_lo_cc(P::mod(5), m, V[5]);
        V[6] = ptx::madc_hi_cc(P::mod(5), m, V[6]);
        V[7] = ptx::madc_lo_cc(P::mod(7), m, V[7]);
        if (st == 0) { V[8] = ptx::madc_hi_cc(P::mod(7), m, V[8]); V[9] = ptx::addc(V[9], 0u); }
        else V[8] = ptx::madc_hi(P::mod(7), m, V[8]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = S[k + 2];
    cond_sub_mod<P>(r);
}

}  // namespace limb

// ---------------------------------------------------------------------------------------------------
template <class P>
struct alignas(16) fp {
    uint32_t l[8];

    TSG_HD static constexpr uint32_t modulus_limb(int i) { return P::mod(i); }
    TSG_HD static fp zero() { fp r; for (int i = 0; i < 8; ++i) r.l[i] = 0; return r; }
    TSG_HD static fp one() { fp r; for (int i = 0; i < 8; ++i) r.l[i] = P::one(i); return r; }
    TSG_HD static fp r2() { fp r; for (int i = 0; i < 8; ++i) r.l[i] = P::r2(i); return r; }
    TSG_HD bool is_zero() const { uint32_t o = 0; for (int i = 0; i < 8; ++i) o |= l[i]; return o == 0; }
    TSG_HD bool operator==(const fp& b) const { uint32_t o = 0; for (int i = 0; i < 8; ++i) o |= l[i] ^ b.l[i]; return o == 0; }
    TSG_HD bool operator!=(const fp& b) const { return !(*this == b); }
    TSG_HD fp operator+(const fp& b) const { fp r; limb::add<P>(r.l, l, b.l); return r; }
    TSG_HD fp operator-(const fp& b) const { fp r; limb::sub<P>(r.l, l, b.l); return r; }
    TSG_HD fp operator*(const fp& b) const { fp r; limb::mont_mul<P>(r.l, l, b.l); return r; }
    // a dedicated squaring (28 doubled off-diagonal products + 8 squares) was measured in round 1: no gain on B200 (profiles/r01_kernel_variants.md)
    TSG_HD fp sqr() const { fp r; limb::mont_mul<P>(r.l, l, l); return r; }
    // a * b - c * d with a single reduction (limb::mont_mul_sub)
    TSG_HD static fp mul_sub(const fp& a, const fp& b, const fp& c, const fp& d) { fp r; limb::mont_mul_sub<P>(r.l, a.l, b.l, c.l, d.l); return r; }
    TSG_HD fp dbl() const { fp r; limb::add<P>(r.l, l, l); return r; }
    TSG_HD fp neg() const { fp z = zero(); fp r; limb::sub<P>(r.l, z.l, l); return r; }
    // Montgomery -> canonical integer limbs
    TSG_HD fp from_mont() const {
        uint32_t T[16];
        for (int i = 0; i < 8; ++i) { T[i] = l[i]; T[8 + i] = 0; }
        fp r; limb::mont_reduce<P>(r.l, T); return r;
    }
    TSG_HD fp to_mont() const { return *this * r2(); }
    TSG_HD static fp from_u64(unsigned long long v) { fp r = zero(); r.l[0] = (uint32_t)v; r.l[1] = (uint32_t)(v >> 32); return r.to_mont(); }
    // a^e for a 64-bit exponent
    TSG_HD fp pow_u64(unsigned long long e) const {
        fp acc = one(); fp base = *this;
        while (e) { if (e & 1) acc = acc * base; base = base.sqr(); e >>= 1; }
        return acc;
    }
    // Fermat inverse a^(p-2); 0 -> 0
    TSG_HD fp inverse() const {
        uint32_t e[8];
        for (int i = 0; i < 8; ++i) e[i] = P::mod(i);
        e[0] -= 2;
        fp acc = one();
        for (int i = 255; i >= 0; --i) {
            acc = acc.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1) acc = acc * *this;
        }
        return acc;
    }
};

typedef fp<FrP> fr_t;
typedef fp<FqP> fq_t;

// multiplier table of a per-launch constant (limb::mul_ctab): passed to kernels by value, so the 64 table limbs are
// constant-bank operands of the multiply-adds and cost no registers
template <class P>
struct fp_ctab {
    uint32_t t[8][8];
    // T[k] = canonical(r * 2^(32 k + 64)); r in Montgomery form
    TSG_HD static fp_ctab make(const fp<P>& r) {
        fp_ctab c;
        const fp<P> w = fp<P>::from_u64(1ull << 32);
        fp<P> cur = r * w * w;
        for (int k = 0; k < 8; ++k) {
            fp<P> can = cur.from_mont();
            for (int i = 0; i < 8; ++i) c.t[k][i] = can.l[i];
            cur = cur * w;
        }
        return c;
    }
    // a * r (Montgomery form in, Montgomery form out): bit-identical to a * r_mont
    TSG_HD fp<P> mul(const fp<P>& a) const { fp<P> r; limb::mul_ctab<P>(r.l, a.l, t); return r; }
};
typedef fp_ctab<FrP> fr_ctab;

// Lazy accumulator for sums of products of reduced operands (carry-chain form, 512 bits).  A carry-free radix-2^29 form (81 plain
// IMAD.WIDE per product) and plain-product reduction rows were built and measured slower in round 1 (ALU-bound; profiles/r01_kernel_variants.md).
template <class P>
struct wide_acc32 {
    uint32_t t[16];
    int pending;   // products added since the last normalisation
    TSG_HD void clear() { for (int i = 0; i < 16; ++i) t[i] = 0; pending = 0; }
    TSG_HD void add_product(const fp<P>& a, const fp<P>& b) {
        uint32_t p[16];
        limb::mul_wide(p, a.l, b.l);
        limb::wide_add(t, p);
        if (++pending == 16) { limb::wide_normalize<P>(t); pending = 0; }
    }
    TSG_HD fp<P> reduce() {
        limb::wide_normalize<P>(t); pending = 0;
        fp<P> r; limb::mont_reduce<P>(r.l, t); return r;
    }
};

template <class P> using wide_acc = wide_acc32<P>;

}  // namespace tsg
