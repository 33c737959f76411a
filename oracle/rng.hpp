// oracle/rng.hpp - ChaCha20Rng, Fp::rand, SipHash-1-3 and the Fiat-Shamir Transcript.
// TEST INFRASTRUCTURE ONLY.
//
// Restates (sources not vendored in /root/reference; versions from its Cargo.lock):
//   rand_chacha 0.3.1 ChaCha20Rng over rand_core 0.6.4 BlockRng (64-word buffer, next_u64 rule),
//   ark-ff 0.4.2 `Fp::rand` (4 x next_u64, mask top 2 bits, rejection; raw limbs ARE the
//   Montgomery representation), Rust std DefaultHasher (SipHash-1-3, keys 0,0),
//   and the reference's own Transcript (src/utils.rs:134-204).
#pragma once
#include <string>
#include <vector>
#include "ff.hpp"

namespace orc {

static inline uint32_t rotl32(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
static inline uint64_t rotl64(uint64_t v, int c) { return (v << c) | (v >> (64 - c)); }

struct ChaCha20Rng {
    uint32_t key[8];
    uint64_t counter;
    uint32_t buf[64];
    int index;

    explicit ChaCha20Rng(const uint8_t seed[32]) : counter(0), index(64) { memcpy(key, seed, 32); }

    static void block(const uint32_t key[8], uint64_t ctr, uint32_t out[16]) {
        uint32_t st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                           key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                           (uint32_t)ctr, (uint32_t)(ctr >> 32), 0, 0};
        uint32_t x[16]; memcpy(x, st, 64);
#define ORC_QR(a, b, c, d)                                     \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16);              \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12);              \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);               \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
        for (int r = 0; r < 10; ++r) {
            ORC_QR(0, 4, 8, 12) ORC_QR(1, 5, 9, 13) ORC_QR(2, 6, 10, 14) ORC_QR(3, 7, 11, 15)
            ORC_QR(0, 5, 10, 15) ORC_QR(1, 6, 11, 12) ORC_QR(2, 7, 8, 13) ORC_QR(3, 4, 9, 14)
        }
#undef ORC_QR
        for (int i = 0; i < 16; ++i) out[i] = x[i] + st[i];
    }
    void generate() {
        for (int b = 0; b < 4; ++b) block(key, counter + b, buf + 16 * b);
        counter += 4;
    }
    uint64_t next_u64() {   // rand_core 0.6.4 BlockRng::next_u64
        if (index < 63) {
            uint64_t v = ((uint64_t)buf[index + 1] << 32) | buf[index];
            index += 2; return v;
        } else if (index >= 64) {
            generate(); index = 2;
            return ((uint64_t)buf[1] << 32) | buf[0];
        } else {
            uint64_t x = buf[63];
            generate(); index = 1;
            return ((uint64_t)buf[0] << 32) | x;
        }
    }
    void fill_bytes(uint8_t* dst, size_t n) {   // whole u32 words consumed, LE
        size_t done = 0;
        while (done < n) {
            if (index >= 64) { generate(); index = 0; }
            uint32_t w = buf[index++];
            for (int k = 0; k < 4 && done < n; ++k) dst[done++] = (uint8_t)(w >> (8 * k));
        }
    }
    // ark-ff Fp::rand for a 254-bit modulus: returns the element whose Montgomery limbs are the draw
    template <class F>
    F rand_fp() {
        for (;;) {
            F t;
            for (int i = 0; i < 4; ++i) t.l[i] = next_u64();
            t.l[3] &= 0xffffffffffffffffull >> 2;
            if (!F::geq_mod(t.l)) return t;
        }
    }
};

// SipHash-1-3, streaming not needed: hash a whole buffer
static inline uint64_t siphash13(const uint8_t* data, size_t n, uint64_t k0 = 0, uint64_t k1 = 0) {
    uint64_t v0 = k0 ^ 0x736f6d6570736575ull, v1 = k1 ^ 0x646f72616e646f6dull;
    uint64_t v2 = k0 ^ 0x6c7967656e657261ull, v3 = k1 ^ 0x7465646279746573ull;
#define ORC_SIPROUND                                                          \
    v0 += v1; v1 = rotl64(v1, 13); v1 ^= v0; v0 = rotl64(v0, 32);             \
    v2 += v3; v3 = rotl64(v3, 16); v3 ^= v2;                                  \
    v0 += v3; v3 = rotl64(v3, 21); v3 ^= v0;                                  \
    v2 += v1; v1 = rotl64(v1, 17); v1 ^= v2; v2 = rotl64(v2, 32);
    size_t full = n & ~(size_t)7;
    for (size_t i = 0; i < full; i += 8) {
        uint64_t m; memcpy(&m, data + i, 8);
        v3 ^= m; ORC_SIPROUND v0 ^= m;
    }
    uint64_t b = (uint64_t)(n & 0xff) << 56;
    for (size_t i = full; i < n; ++i) b |= (uint64_t)data[i] << (8 * (i - full));
    v3 ^= b; ORC_SIPROUND v0 ^= b;
    v2 ^= 0xff;
    ORC_SIPROUND ORC_SIPROUND ORC_SIPROUND
#undef ORC_SIPROUND
    return v0 ^ v1 ^ v2 ^ v3;
}

// src/utils.rs:134-204
struct Transcript {
    std::vector<uint8_t> state;   // never reset; challenge labels are appended too (:173)

    void append_bytes(const void* p, size_t n) {
        const uint8_t* b = (const uint8_t*)p;
        state.insert(state.end(), b, b + n);
    }
    void append_field_element(const std::string& label, const Fr& x) {      // :150-158
        append_bytes(label.data(), label.size());
        uint8_t b[32]; x.to_bytes_le(b); append_bytes(b, 32);
    }
    void append_field_elements(const std::string& label, const Fr* xs, size_t n) {   // :161-169
        append_bytes(label.data(), label.size());
        for (size_t i = 0; i < n; ++i) { uint8_t b[32]; xs[i].to_bytes_le(b); append_bytes(b, 32); }
    }
    Fr challenge_field_element(const std::string& label) {                  // :172-192
        append_bytes(label.data(), label.size());
        // Vec<u8>::hash: length prefix (usize LE) then the bytes, through SipHash-1-3(0,0)
        std::vector<uint8_t> msg(8 + state.size());
        uint64_t len = state.size();
        memcpy(msg.data(), &len, 8);
        memcpy(msg.data() + 8, state.data(), state.size());
        uint64_t h = siphash13(msg.data(), msg.size());
        uint8_t seed[32];
        for (int i = 0; i < 4; ++i) memcpy(seed + 8 * i, &h, 8);
        ChaCha20Rng rng(seed);
        return rng.rand_fp<Fr>();
    }
    std::vector<Fr> challenge_field_elements(const std::string& label, size_t count) {   // :195-203
        std::vector<Fr> out;
        for (size_t i = 0; i < count; ++i) out.push_back(challenge_field_element(label + "_" + std::to_string(i)));
        return out;
    }
};

}  // namespace orc
