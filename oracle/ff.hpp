// oracle/ff.hpp - CPU restatement of the BN254 field arithmetic the reference leans on.
//
// TEST INFRASTRUCTURE ONLY (see oracle/README.md): the product never links this.
//
// Restates ark-ff 0.4.2 `Fp<MontBackend<_,4>,4>`: one BigInt<4> = uint64_t[4] little-endian
// limbs holding a*2^256 mod p (reference alias: src/utils.rs:14 `FieldElement = ark_bn254::Fr`).
// ark-ff is not vendored under /root/reference (Cargo.lock pins ark-ff 0.4.2, no `asm`
// feature -> portable 4x64 CIOS); the algorithm below is the published CIOS Montgomery product.
#pragma once
#include <cstdint>
#include <cstring>
#include <cstddef>

namespace orc {

typedef unsigned __int128 u128;

struct FrParams {
    static constexpr uint64_t MOD[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    static constexpr uint64_t ONE[4] = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full};
    static constexpr uint64_t R2[4]  = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
    static constexpr uint64_t INV = 0xc2e1f593efffffffull;
};
struct FqParams {
    static constexpr uint64_t MOD[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    static constexpr uint64_t ONE[4] = {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full};
    static constexpr uint64_t R2[4]  = {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full};
    static constexpr uint64_t INV = 0x87d20782e4866389ull;
};

template <class P>
struct Fp {
    uint64_t l[4];

    static Fp zero() { Fp r; r.l[0] = r.l[1] = r.l[2] = r.l[3] = 0; return r; }
    static Fp one() { Fp r; memcpy(r.l, P::ONE, 32); return r; }
    static Fp from_raw(const uint64_t* p) { Fp r; memcpy(r.l, p, 32); return r; }
    // canonical small integer -> Montgomery
    static Fp from_u64(uint64_t v) {
        Fp r; r.l[0] = v; r.l[1] = r.l[2] = r.l[3] = 0;
        Fp r2; memcpy(r2.l, P::R2, 32);
        return r * r2;
    }
    // canonical 4-limb integer (< 2^256, any value) -> Montgomery of (v mod p)
    static Fp from_canonical_limbs(const uint64_t* v) {
        Fp r; memcpy(r.l, v, 32);
        // reduce below p by repeated subtraction (2^256 < 6p)
        while (geq_mod(r.l)) sub_mod_raw(r.l);
        Fp r2; memcpy(r2.l, P::R2, 32);
        return r * r2;
    }
    void to_canonical_limbs(uint64_t* out) const {
        Fp o; o.l[0] = 1; o.l[1] = o.l[2] = o.l[3] = 0;   // raw 1 -> multiplying strips one R
        Fp c = (*this) * o;
        memcpy(out, c.l, 32);
    }
    void to_bytes_le(uint8_t* out) const {   // ark-serialize: canonical integer, 32 bytes LE
        uint64_t c[4]; to_canonical_limbs(c);
        memcpy(out, c, 32);
    }

    static bool geq_mod(const uint64_t* a) {
        for (int i = 3; i >= 0; --i) {
            if (a[i] > P::MOD[i]) return true;
            if (a[i] < P::MOD[i]) return false;
        }
        return true;
    }
    static void sub_mod_raw(uint64_t* a) {
        u128 borrow = 0;
        for (int i = 0; i < 4; ++i) {
            u128 d = (u128)a[i] - P::MOD[i] - borrow;
            a[i] = (uint64_t)d;
            borrow = (d >> 64) & 1;
        }
    }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
    bool operator==(const Fp& o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }
    bool operator!=(const Fp& o) const { return !(*this == o); }

    Fp operator+(const Fp& o) const {
        Fp r; u128 c = 0;
        for (int i = 0; i < 4; ++i) { c += (u128)l[i] + o.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
        if (c || geq_mod(r.l)) sub_mod_raw(r.l);
        return r;
    }
    Fp operator-(const Fp& o) const {
        Fp r; u128 borrow = 0;
        for (int i = 0; i < 4; ++i) {
            u128 d = (u128)l[i] - o.l[i] - borrow;
            r.l[i] = (uint64_t)d; borrow = (d >> 64) & 1;
        }
        if (borrow) {
            u128 c = 0;
            for (int i = 0; i < 4; ++i) { c += (u128)r.l[i] + P::MOD[i]; r.l[i] = (uint64_t)c; c >>= 64; }
        }
        return r;
    }
    Fp neg() const { return is_zero() ? *this : zero() - *this; }
    Fp dbl() const { return *this + *this; }

    // CIOS Montgomery product
    Fp operator*(const Fp& o) const {
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            u128 c = 0;
            for (int j = 0; j < 4; ++j) {
                c += (u128)l[j] * o.l[i] + t[j];
                t[j] = (uint64_t)c; c >>= 64;
            }
            c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * P::INV;
            c = (u128)m * P::MOD[0] + t[0]; c >>= 64;
            for (int j = 1; j < 4; ++j) {
                c += (u128)m * P::MOD[j] + t[j];
                t[j - 1] = (uint64_t)c; c >>= 64;
            }
            c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
        }
        Fp r; memcpy(r.l, t, 32);
        if (t[4] || geq_mod(r.l)) sub_mod_raw(r.l);
        return r;
    }
    Fp& operator+=(const Fp& o) { *this = *this + o; return *this; }
    Fp& operator-=(const Fp& o) { *this = *this - o; return *this; }
    Fp& operator*=(const Fp& o) { *this = *this * o; return *this; }
    Fp sqr() const { return *this * *this; }

    Fp pow_limbs(const uint64_t* e, int nl) const {
        Fp acc = one();
        bool started = false;
        for (int i = nl * 64 - 1; i >= 0; --i) {
            if (started) acc = acc.sqr();
            if ((e[i / 64] >> (i % 64)) & 1) { acc = acc * *this; started = true; }
        }
        return acc;
    }
    Fp pow_u64(uint64_t e) const { return pow_limbs(&e, 1); }
    // Fermat inverse; inverse of zero is zero (callers check)
    Fp inverse() const {
        uint64_t e[4]; memcpy(e, P::MOD, 32);
        e[0] -= 2;   // MOD[0] low limb >= 2 for both fields
        return pow_limbs(e, 4);
    }
};

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

}  // namespace orc
