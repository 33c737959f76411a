// host/statement_digest.hpp - binding digest of a clear-text statement for the non-parity constraint sum-checks
// (host/read_check.cpp, host/memory_check.cpp).  Fiat-Shamir needs every challenge to depend on the statement: the
// digest is absorbed into the transcript BEFORE the first challenge is drawn, by prover and verifier alike.
//
// The statement can be hundreds of MiB (BASELINE config 3: 2^22 lookups), and the reference Transcript re-hashes its
// whole state at every challenge (src/utils.rs:172-192), so the statement itself cannot live in the transcript.  It is
// reduced to 32 bytes by a two-level BLAKE2b-256 tree (RFC 7693, unkeyed, sequential mode per node):
//
//   leaf(segment s, chunk c) = BLAKE2b-256( bytes [c * 2^20, min((c + 1) * 2^20, len_s)) of segment s )
//   root = BLAKE2b-256( domain (16 bytes, zero padded) | u64 #header | header u64s (LE) | u64 #segments |
//                       per segment: u64 byte length | its leaf digests in order )
//
// Leaves are hashed on all host threads.  The root enters the transcript as two field elements (the low and the high
// 128 bits, little-endian integers).
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace tsg { namespace host {

class Blake2b256 {
    uint64_t h_[8], t_ = 0;
    uint8_t buf_[128];
    size_t fill_ = 0;
    static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    static uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }   // little-endian hosts only (x86-64, aarch64)
    static const uint64_t* iv() {
        static const uint64_t v[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                      0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        return v;
    }
    void compress(const uint8_t* block, bool last) {
        static const uint8_t S[12][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
        uint64_t m[16], v[16];
        for (int i = 0; i < 16; ++i) m[i] = load64(block + 8 * i);
        for (int i = 0; i < 8; ++i) { v[i] = h_[i]; v[i + 8] = iv()[i]; }
        v[12] ^= t_;                       // messages stay far below 2^64 bytes: the high counter word is zero
        if (last) v[14] = ~v[14];
#define TSG_B2_G(a, b, c, d, x, y) \
        v[a] += v[b] + (x); v[d] = rotr(v[d] ^ v[a], 32); v[c] += v[d]; v[b] = rotr(v[b] ^ v[c], 24); \
        v[a] += v[b] + (y); v[d] = rotr(v[d] ^ v[a], 16); v[c] += v[d]; v[b] = rotr(v[b] ^ v[c], 63);
        for (int r = 0; r < 12; ++r) {
            const uint8_t* s = S[r];
            TSG_B2_G(0, 4, 8, 12, m[s[0]], m[s[1]])   TSG_B2_G(1, 5, 9, 13, m[s[2]], m[s[3]])
            TSG_B2_G(2, 6, 10, 14, m[s[4]], m[s[5]])  TSG_B2_G(3, 7, 11, 15, m[s[6]], m[s[7]])
            TSG_B2_G(0, 5, 10, 15, m[s[8]], m[s[9]])  TSG_B2_G(1, 6, 11, 12, m[s[10]], m[s[11]])
            TSG_B2_G(2, 7, 8, 13, m[s[12]], m[s[13]]) TSG_B2_G(3, 4, 9, 14, m[s[14]], m[s[15]])
        }
#undef TSG_B2_G
        for (int i = 0; i < 8; ++i) h_[i] ^= v[i] ^ v[i + 8];
    }

public:
    Blake2b256() {
        for (int i = 0; i < 8; ++i) h_[i] = iv()[i];
        h_[0] ^= 0x01010000ULL ^ 32;       // digest length 32, no key, fanout 1, depth 1
    }
    void update(const void* data, size_t n) {
        const uint8_t* p = (const uint8_t*)data;
        while (n) {
            if (fill_ == 128) { t_ += 128; compress(buf_, false); fill_ = 0; }   // a full buffer is only compressed once more input follows
            size_t take = std::min(n, (size_t)128 - fill_);
            memcpy(buf_ + fill_, p, take);
            fill_ += take; p += take; n -= take;
        }
    }
    void update_u64(uint64_t v) { update(&v, 8); }
    void final(uint8_t out[32]) {
        t_ += fill_;
        memset(buf_ + fill_, 0, 128 - fill_);
        compress(buf_, true);
        memcpy(out, h_, 32);
    }
};

struct StatementSegment { const void* data; size_t bytes; };

inline void statement_digest(const char* domain, const uint64_t* header, size_t num_header, const StatementSegment* segs, size_t num_segs,
                             uint8_t out[32]) {
    const size_t CH = (size_t)1 << 20;
    struct Leaf { size_t seg, off, len; };
    std::vector<Leaf> leaves;
    for (size_t s = 0; s < num_segs; ++s)
        for (size_t off = 0; off < segs[s].bytes; off += CH) leaves.push_back({s, off, std::min(CH, segs[s].bytes - off)});
    std::vector<uint8_t> dig(32 * (leaves.size() ? leaves.size() : 1));
    auto work = [&](std::atomic<size_t>* next) {
        for (size_t i; (i = next->fetch_add(1)) < leaves.size();) {
            Blake2b256 b;
            b.update((const uint8_t*)segs[leaves[i].seg].data + leaves[i].off, leaves[i].len);
            b.final(&dig[32 * i]);
        }
    };
    std::atomic<size_t> next(0);
    unsigned nt = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), std::min<size_t>(32, leaves.size()));
    if (nt <= 1) work(&next);
    else {
        std::vector<std::thread> th;
        for (unsigned i = 0; i < nt; ++i) th.emplace_back(work, &next);
        for (auto& t : th) t.join();
    }
    Blake2b256 root;
    uint8_t dom[16] = {0};
    memcpy(dom, domain, std::min<size_t>(16, strlen(domain)));
    root.update(dom, 16);
    root.update_u64(num_header);
    for (size_t i = 0; i < num_header; ++i) root.update_u64(header[i]);
    root.update_u64(num_segs);
    size_t li = 0;
    for (size_t s = 0; s < num_segs; ++s) {
        root.update_u64(segs[s].bytes);
        for (size_t off = 0; off < segs[s].bytes; off += CH, ++li) root.update(&dig[32 * li], 32);
    }
    root.final(out);
}

}}  // namespace tsg::host
