// sumcheck.cu - sum-check prover round kernels over BN254 Fr tables resident in HBM.
//
// Replaces the reference's closure-driven round (SumCheck::compute_round_polynomial,
// src/sumcheck.rs:156-207, which re-evaluates whole MLEs via src/polynomials.rs:85-122 for each of
// 4 * 2^remaining points) by the table form: round k pairs the entries that differ in variable k,
// g(X) = sum_pairs prod_t (lo_t + X (hi_t - lo_t)) at X = 0..3, then binds
// T'[.] = lo + r (hi - lo) (what MultilinearExtension::partial_evaluate computes, polynomials.rs:126-161).
// The values are exact field elements, so they equal the reference's.
//
// HBM layout: a table of 2^n entries is stored in BIT-REVERSED index order (entry with reference
// index i lives at position bitrev_n(i)); 32 bytes per entry, reference Montgomery limbs.  Variable k
// (bit k of the reference index, polynomials.rs:111-118) is then the TOP bit of the position in round
// k, so a round streams two (bind) or four (fused bind + next-round evaluation) perfectly contiguous
// ranges with one 256-bit load per entry per thread, and the bound table is written in place to the
// low half - again contiguous.  Algorithmic bytes: evaluation 32 d N_k, bind 48 d N_k.
#include "fr_device.cuh"
#include "sumcheck.cuh"
#include "tma_stream.cuh"
#include "../host/field64.hpp"
#include <atomic>
#include <chrono>
#include <cstring>

namespace tsg {

// ---------------------------------------------------------------- round sums over the ranks, inside the round kernel
// Sharded SumCheck::prove (comm.cu): every rank sums its slice of the hypercube; the round values are the sums over the ranks.  Instead of
// a separate widen kernel + ncclAllReduce + copy per round, the first warp of the FINISHING block of the round kernel exchanges the block-reduced
// sums with its peers directly (lane g serves peer g): it stores its NV field elements into slot [parity][rank] of every peer's mailbox (peer-mapped device memory: plain
// stores that travel over NVLink / NVSwitch), publishes them with a system-scope fence + sequence flag, spins until the G slots of its
// own mailbox carry the same sequence number, and adds them in rank order (exact modular additions: every rank gets the same bits).  The
// epilogue then writes the global values to the pinned host mirror as on one GPU.  Two parities: a fast peer may already deliver round
// k + 1 while this rank still reads round k (it cannot get further ahead: round k + 1 needs this rank's own contribution).
// State lives in device globals of this translation unit (the kernels are here); comm.cu sets it through sc_peer_*.
struct PeerState {
    unsigned char* mbox[SC_MAX_PEERS];   // mailbox of every rank as mapped into THIS process (own rank: the local allocation)
    int* host_err;                       // pinned: set to 1 when a peer did not answer within the time limit
    int nranks, rank;
    unsigned on;                         // exchange enabled for the launches that follow (sharded prove in progress)
    unsigned seq;                        // exchanges done so far (all ranks run the same sequence of launches)
};
__device__ PeerState g_peer;
constexpr unsigned long long PEER_TIMEOUT_CYCLES = 6000000000ull;   // ~3 s at 2 GHz: a rank that died must not hang the others' GPUs

// 64-bit volatile accesses for everything that crosses the link or is written by a peer (no L1, no wide vector forms on peer mappings)
__device__ __forceinline__ void peer_store(fr_t* dst, const fr_t& v) {
    volatile unsigned long long* d = (volatile unsigned long long*)dst;
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i] = (unsigned long long)v.l[2 * i] | ((unsigned long long)v.l[2 * i + 1] << 32);
}
__device__ __forceinline__ fr_t peer_load(const fr_t* src) {
    const volatile unsigned long long* s = (const volatile unsigned long long*)src;
    fr_t r;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const unsigned long long w = s[i]; r.l[2 * i] = (uint32_t)w; r.l[2 * i + 1] = (uint32_t)(w >> 32); }
    return r;
}

// run by the 32 lanes of warp 0 of the finishing block: v valid in lane 0 on entry and on exit; lane g serves peer g
template <int NV>
__device__ __noinline__ void peer_sum(fr_t (&v)[NV], fr_t* smem) {
    static_assert(NV * sizeof(fr_t) <= SC_PEER_SLOT_DATA, "slot too small");
    const int G = g_peer.nranks, me = g_peer.rank;
    const int lane = threadIdx.x & 31;
    const unsigned seq = g_peer.seq + 1;
    const size_t par = (size_t)(seq & 1u) * SC_MAX_PEERS * SC_PEER_SLOT;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) smem[k] = v[k];
    }
    __syncwarp();
    const unsigned char* mine = g_peer.mbox[me] + par;
    if (lane < G) {
        unsigned char* slot = g_peer.mbox[lane] + par + (size_t)me * SC_PEER_SLOT;
#pragma unroll
        for (int k = 0; k < NV; ++k) peer_store((fr_t*)slot + k, smem[k]);
        __threadfence_system();
        *(volatile unsigned*)(slot + SC_PEER_SLOT_DATA) = seq;
        const volatile unsigned* flag = (const volatile unsigned*)(mine + (size_t)lane * SC_PEER_SLOT + SC_PEER_SLOT_DATA);
        const long long t0 = clock64();
        while (*flag != seq) {
            if ((unsigned long long)(clock64() - t0) > PEER_TIMEOUT_CYCLES) { *g_peer.host_err = 1; break; }
        }
        __threadfence_system();
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] = fr_t::zero();
        for (int g = 0; g < G; ++g) {                       // rank order: every rank adds the same values in the same order
            const fr_t* src = (const fr_t*)(mine + (size_t)g * SC_PEER_SLOT);
#pragma unroll
            for (int k = 0; k < NV; ++k) v[k] = v[k] + peer_load(src + k);
        }
        g_peer.seq = seq;
    }
}

cudaError_t sc_peer_configure(unsigned char* const* mbox, int nranks, int rank, int* host_err, cudaStream_t s) {
    PeerState st;
    memset(&st, 0, sizeof(st));
    for (int g = 0; g < nranks && g < SC_MAX_PEERS; ++g) st.mbox[g] = mbox[g];
    st.host_err = host_err; st.nranks = nranks; st.rank = rank; st.on = 0; st.seq = 0;
    cudaError_t e = cudaMemcpyToSymbolAsync(g_peer, &st, sizeof(st), 0, cudaMemcpyHostToDevice, s);
    return e ? e : cudaStreamSynchronize(s);
}
__global__ void k_peer_switch(unsigned on) { g_peer.on = on; }
cudaError_t sc_peer_enable(bool on, cudaStream_t s) { k_peer_switch<<<1, 1, 0, s>>>(on ? 1u : 0u); return cudaGetLastError(); }

// ---------------------------------------------------------------- small all-gather over the same mailboxes (one block)
// in: `bytes` (multiple of 8, <= SC_PEER_AG_DATA) of this rank, readable by the device (pinned host memory or HBM); out: nranks * bytes, rank-major
// (pinned host memory: the host reads it after the stream synchronisation).  Region 2 of the mailbox, own sequence counter.
__device__ unsigned g_peer_ag_seq;
__global__ void __launch_bounds__(64) k_peer_allgather(const unsigned long long* in, unsigned words, unsigned long long* out) {
    const int G = g_peer.nranks, me = g_peer.rank;
    const unsigned seq = g_peer_ag_seq + 1;
    const size_t base = (size_t)2 * SC_MAX_PEERS * SC_PEER_SLOT + (size_t)(seq & 1u) * SC_MAX_PEERS * SC_PEER_AG_SLOT;
    const unsigned t = threadIdx.x;
    if (t < words) {
        const unsigned long long w = in[t];
        for (int g = 0; g < G; ++g) ((unsigned long long*)(g_peer.mbox[g] + base + (size_t)me * SC_PEER_AG_SLOT))[t] = w;
    }
    __threadfence_system();
    __syncthreads();
    if ((int)t < G) {
        *(volatile unsigned*)(g_peer.mbox[t] + base + (size_t)me * SC_PEER_AG_SLOT + SC_PEER_AG_DATA) = seq;
        const volatile unsigned* flag = (const volatile unsigned*)(g_peer.mbox[me] + base + (size_t)t * SC_PEER_AG_SLOT + SC_PEER_AG_DATA);
        const long long t0 = clock64();
        while (*flag != seq) {
            if ((unsigned long long)(clock64() - t0) > PEER_TIMEOUT_CYCLES) { *g_peer.host_err = 1; break; }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (t < words)
        for (int g = 0; g < G; ++g)
            out[(size_t)g * words + t] = ((const volatile unsigned long long*)(g_peer.mbox[me] + base + (size_t)g * SC_PEER_AG_SLOT))[t];
    if (t == 0) g_peer_ag_seq = seq;
}
cudaError_t launch_peer_allgather(const void* in, size_t bytes, void* out, cudaStream_t s) {
    if (bytes % 8 || bytes > SC_PEER_AG_DATA) return cudaErrorInvalidValue;
    k_peer_allgather<<<1, 64, 0, s>>>((const unsigned long long*)in, (unsigned)(bytes / 8), (unsigned long long*)out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- per-pair evaluation contributions
template <int D> struct EvalAcc;

// d = 1: g is linear; only g(0), g(1) are accumulated, g(2), g(3) follow by extrapolation.
template <> struct EvalAcc<1> {
    static constexpr int NV = 2;
    fr_t e0, e1;
    __device__ __forceinline__ void clear() { e0 = fr_t::zero(); e1 = fr_t::zero(); }
    __device__ __forceinline__ void pair(const fr_t* lo, const fr_t* hi) { e0 = e0 + lo[0]; e1 = e1 + hi[0]; }
    __device__ __forceinline__ void finish(fr_t (&v)[NV]) { v[0] = e0; v[1] = e1; }
    __device__ static void expand(const fr_t (&v)[NV], fr_t* out4) {
        out4[0] = v[0]; out4[1] = v[1];
        fr_t d = v[1] - v[0];
        out4[2] = v[1] + d; out4[3] = out4[2] + d;
    }
};

// d = 2: g is quadratic; g(0), g(1), g(2) as lazily reduced 512-bit dot products, g(3) = g0 - 3 g1 + 3 g2.
template <> struct EvalAcc<2> {
    static constexpr int NV = 3;
    wide_acc<FrP> a0, a1, a2;
    __device__ __forceinline__ void clear() { a0.clear(); a1.clear(); a2.clear(); }
    __device__ __forceinline__ void pair(const fr_t* lo, const fr_t* hi) {
        a0.add_product(lo[0], lo[1]);
        a1.add_product(hi[0], hi[1]);
        fr_t x = hi[0] + (hi[0] - lo[0]);
        fr_t y = hi[1] + (hi[1] - lo[1]);
        a2.add_product(x, y);
    }
    __device__ __forceinline__ void finish(fr_t (&v)[NV]) { v[0] = a0.reduce(); v[1] = a1.reduce(); v[2] = a2.reduce(); }
    __device__ static void expand(const fr_t (&v)[NV], fr_t* out4) {
        out4[0] = v[0]; out4[1] = v[1]; out4[2] = v[2];
        fr_t d = v[2] - v[1];
        out4[3] = v[0] + d + d + d;
    }
};

// d = 2 when the round's claim is known: g(0) + g(1) = claim (the identity SumCheck::prove checks, src/sumcheck.rs:77-84, which
// holds by construction once round 0 has been checked in full), so only g(0) and g(2) are accumulated: two lazy products per pair.
struct EvalAcc2Claim {
    static constexpr int NV = 2;
    wide_acc<FrP> a0, a2;
    __device__ __forceinline__ void clear() { a0.clear(); a2.clear(); }
    __device__ __forceinline__ void pair(const fr_t* lo, const fr_t* hi) {
        a0.add_product(lo[0], lo[1]);
        fr_t x = hi[0] + (hi[0] - lo[0]);
        fr_t y = hi[1] + (hi[1] - lo[1]);
        a2.add_product(x, y);
    }
    __device__ __forceinline__ void finish(fr_t (&v)[NV]) { v[0] = a0.reduce(); v[1] = a2.reduce(); }
};
struct EvalClaimEpilogue {
    fr_t* out4; fr_t claim;
    __device__ void warp(fr_t (&v)[2], fr_t* smem) const {
        if (g_peer.on) peer_sum<2>(v, smem);           // sharded: v = sums over all ranks; the host derives g(1) from the GLOBAL claim
        if ((threadIdx.x & 31) == 0) (*this)(v);
    }
    __device__ void operator()(fr_t (&v)[2]) const {
        fr_t g1 = claim - v[0];
        fr_t d = v[1] - g1;
        out4[0] = v[0]; out4[1] = g1; out4[2] = v[1]; out4[3] = v[0] + d + d + d;
    }
};

// d = 3: cubic; all four points, one Montgomery product + one lazy product per point.
template <> struct EvalAcc<3> {
    static constexpr int NV = 4;
    wide_acc<FrP> a[4];
    __device__ __forceinline__ void clear() { for (int i = 0; i < 4; ++i) a[i].clear(); }
    __device__ __forceinline__ void pair(const fr_t* lo, const fr_t* hi) {
        fr_t v0 = lo[0], v1 = lo[1], v2 = lo[2];
        fr_t d0 = hi[0] - lo[0], d1 = hi[1] - lo[1], d2 = hi[2] - lo[2];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            a[x].add_product(v0 * v1, v2);
            v0 = v0 + d0; v1 = v1 + d1; v2 = v2 + d2;
        }
    }
    __device__ __forceinline__ void finish(fr_t (&v)[NV]) { for (int i = 0; i < 4; ++i) v[i] = a[i].reduce(); }
    __device__ static void expand(const fr_t (&v)[NV], fr_t* out4) { for (int i = 0; i < 4; ++i) out4[i] = v[i]; }
};

template <int D>
struct EvalEpilogue {
    fr_t* out4;
    __device__ void warp(fr_t (&v)[EvalAcc<D>::NV], fr_t* smem) const {
        if (g_peer.on) peer_sum<EvalAcc<D>::NV>(v, smem);   // sharded: sums over all ranks (extrapolation is linear, so it commutes with the sum)
        if ((threadIdx.x & 31) == 0) (*this)(v);
    }
    __device__ void operator()(fr_t (&v)[EvalAcc<D>::NV]) const {
        fr_t o[4];
        EvalAcc<D>::expand(v, o);
        for (int i = 0; i < 4; ++i) out4[i] = o[i];
    }
};

// ---------------------------------------------------------------- K2: round evaluation
template <int D>
__global__ void __launch_bounds__(SC_THREADS, (D <= 2 ? 2 : 1)) k_round_eval(ScTables tabs, size_t half, fr_t* partials, unsigned int* ticket, fr_t* out4) {
    __shared__ fr_t smem[EvalAcc<D>::NV * 32];
    EvalAcc<D> acc; acc.clear();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < half; p += stride) {
        fr_t lo[D], hi[D];
#pragma unroll
        for (int t = 0; t < D; ++t) { lo[t] = ld256_stream(tabs.t[t] + p); hi[t] = ld256_stream(tabs.t[t] + p + half); }
        acc.pair(lo, hi);
    }
    fr_t v[EvalAcc<D>::NV];
    acc.finish(v);
    grid_finish_sum<fr_t, EvalAcc<D>::NV>(v, partials, ticket, smem, EvalEpilogue<D>{out4});
}

// d = 2 evaluation with the claim g(0) + g(1) known: sums g(0) and g(2) only (two lazy products per pair instead of three)
__global__ void __launch_bounds__(SC_THREADS, 2) k_round_eval2_claim(ScTables tabs, size_t half, const fr_t claim, fr_t* partials, unsigned int* ticket, fr_t* out4) {
    __shared__ fr_t smem[2 * 32];
    EvalAcc2Claim acc; acc.clear();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < half; p += stride) {
        fr_t lo[2], hi[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) { lo[t] = ld256_stream(tabs.t[t] + p); hi[t] = ld256_stream(tabs.t[t] + p + half); }
        acc.pair(lo, hi);
    }
    fr_t v[2];
    acc.finish(v);
    grid_finish_sum<fr_t, 2>(v, partials, ticket, smem, EvalClaimEpilogue{out4, claim});
}

// ---------------------------------------------------------------- K1: bind (fold) one table in place
__global__ void __launch_bounds__(SC_THREADS) k_bind(fr_t* t, size_t half, const fr_ctab r) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < half; p += stride) {
        fr_t lo = ld256_stream(t + p), hi = ld256_stream(t + p + half);
        st256(t + p, lo + r.mul(hi - lo));
    }
}
// out-of-place variant (keeps the source table intact; used by partial_evaluate on borrowed tables)
__global__ void __launch_bounds__(SC_THREADS) k_bind_to(const fr_t* t, fr_t* out, size_t half, const fr_ctab r) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < half; p += stride) {
        fr_t lo = ld256_nc(t + p), hi = ld256_nc(t + p + half);
        st256(out + p, lo + r.mul(hi - lo));
    }
}

// ---------------------------------------------------------------- K3: fused bind(r_k) + evaluate(round k+1)
// Thread p < quarter owns positions p, p+q, p+2q, p+3q of each table: top position bit = variable k
// (bound now), next bit = variable k+1 (evaluated now).  Writes the bound table to positions [0, 2q).
template <int D>
__global__ void __launch_bounds__(SC_THREADS, (D <= 2 ? 2 : 1)) k_bind_eval(ScTables tabs, size_t quarter, const fr_ctab r, fr_t* partials,
                                                          unsigned int* ticket, fr_t* out4) {
    __shared__ fr_t smem[EvalAcc<D>::NV * 32];
    EvalAcc<D> acc; acc.clear();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < quarter; p += stride) {
        fr_t lo[D], hi[D];
#pragma unroll
        for (int t = 0; t < D; ++t) {
            fr_t* base = tabs.t[t];
            fr_t a0 = ld256_stream(base + p), a1 = ld256_stream(base + p + quarter);
            fr_t b0 = ld256_stream(base + p + 2 * quarter), b1 = ld256_stream(base + p + 3 * quarter);
            lo[t] = a0 + r.mul(b0 - a0);
            hi[t] = a1 + r.mul(b1 - a1);
            st256(base + p, lo[t]);
            st256(base + p + quarter, hi[t]);
        }
        acc.pair(lo, hi);
    }
    fr_t v[EvalAcc<D>::NV];
    acc.finish(v);
    grid_finish_sum<fr_t, EvalAcc<D>::NV>(v, partials, ticket, smem, EvalEpilogue<D>{out4});
}


// d = 2 with the claim of the round being evaluated (= g_k(r), known to the host before the launch)
__global__ void __launch_bounds__(SC_THREADS, 2) k_bind_eval2_claim(ScTables tabs, size_t quarter, const fr_ctab r, const fr_t claim, fr_t* partials,
                                                                   unsigned int* ticket, fr_t* out4) {
    __shared__ fr_t smem[2 * 32];
    EvalAcc2Claim acc; acc.clear();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < quarter; p += stride) {
        fr_t lo[2], hi[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            fr_t* base = tabs.t[t];
            fr_t a0 = ld256_stream(base + p), a1 = ld256_stream(base + p + quarter);
            fr_t b0 = ld256_stream(base + p + 2 * quarter), b1 = ld256_stream(base + p + 3 * quarter);
            lo[t] = a0 + r.mul(b0 - a0);
            hi[t] = a1 + r.mul(b1 - a1);
            st256(base + p, lo[t]);
            st256(base + p + quarter, hi[t]);
        }
        acc.pair(lo, hi);
    }
    fr_t v[2];
    acc.finish(v);
    grid_finish_sum<fr_t, 2>(v, partials, ticket, smem, EvalClaimEpilogue{out4, claim});
}

// ---------------------------------------------------------------- warp-private prefetch variants
// The ncu capture of the plain kernels shows the FMA-heavy pipe 67-69% busy with `long_scoreboard` (waiting for the loads at the top of
// the loop) as the largest stall: 16 warps per SM cannot both hide HBM latency and keep the multiplier fed, and the 64 registers that a
// register-level prefetch of the next eight elements would need do not exist (128 per thread already).  Here the NEXT iteration's data
// waits in shared memory instead: every warp owns one slot (rows x 32 lanes x 32 bytes) and one mbarrier; lane 0 posts one 1 KiB bulk
// copy (cp.async.bulk, the TMA engine) per row for the warp's next tile right after the warp has pulled the current tile into registers,
// so a whole iteration of arithmetic (~2000 cycles) covers the copy.  No CTA-wide barrier, the same 2 x 256 threads per SM and the same
// register budget as the plain kernels; 64 KiB (fused) / 32 KiB (evaluation) of dynamic shared memory per CTA.
constexpr int PF_WARPS = SC_THREADS / 32;

template <int ROWS>
struct WarpPrefetch {
    unsigned char* slot; uint64_t* bar; uint32_t phase; unsigned lane;
    __device__ __forceinline__ void init(unsigned char* dyn, uint64_t* bars) {
        const unsigned warp = threadIdx.x >> 5; lane = threadIdx.x & 31;
        slot = dyn + (size_t)warp * ROWS * 1024; bar = bars + warp; phase = 0;
        if (lane == 0) { tma::mbar_init(bar, 1); tma::fence_barrier_init(); }
        __syncwarp();
    }
    // rows[i] = address of this warp's 32 consecutive elements of stream i
    __device__ __forceinline__ void issue(const fr_t* const (&rows)[ROWS]) {
        if (lane == 0) {
            tma::mbar_expect_tx(bar, ROWS * 1024u);
#pragma unroll
            for (int i = 0; i < ROWS; ++i) tma::bulk_g2s(slot + i * 1024, rows[i], 1024u, bar);
        }
    }
    __device__ __forceinline__ void fetch(fr_t (&e)[ROWS]) {
        tma::mbar_wait(bar, phase); phase ^= 1;
#pragma unroll
        for (int i = 0; i < ROWS; ++i) e[i] = ((const fr_t*)(slot + i * 1024))[lane];
        __syncwarp();          // every lane has read the slot before lane 0 lets the copy engine overwrite it
    }
};

__global__ void __launch_bounds__(SC_THREADS, 2) k_bind_eval2_claim_pf(ScTables tabs, size_t quarter, const fr_ctab r, const fr_t claim, fr_t* partials,
                                                                        unsigned int* ticket, fr_t* out4) {
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ fr_t smem[2 * 32];
    __shared__ uint64_t bars[PF_WARPS];
    WarpPrefetch<8> pf; pf.init(dyn, bars);
    EvalAcc2Claim acc; acc.clear();
    const size_t tiles = quarter / 32, nw = (size_t)gridDim.x * PF_WARPS;
    size_t tile = (size_t)blockIdx.x * PF_WARPS + (threadIdx.x >> 5);
    auto post = [&](size_t t) {
        const fr_t* rows[8];
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) rows[4 * k + j] = tabs.t[k] + j * quarter + t * 32;
        pf.issue(rows);
    };
    if (tile < tiles) post(tile);
    for (; tile < tiles; tile += nw) {
        fr_t e[8];
        pf.fetch(e);
        if (tile + nw < tiles) post(tile + nw);
        const size_t p = tile * 32 + pf.lane;
        fr_t lo[2], hi[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            lo[t] = e[4 * t + 0] + r.mul(e[4 * t + 2] - e[4 * t + 0]);
            hi[t] = e[4 * t + 1] + r.mul(e[4 * t + 3] - e[4 * t + 1]);
            st256(tabs.t[t] + p, lo[t]);
            st256(tabs.t[t] + p + quarter, hi[t]);
        }
        acc.pair(lo, hi);
    }
    fr_t v[2];
    acc.finish(v);
    grid_finish_sum<fr_t, 2>(v, partials, ticket, smem, EvalClaimEpilogue{out4, claim});
}

__global__ void __launch_bounds__(SC_THREADS, 2) k_round_eval2_pf(ScTables tabs, size_t half, fr_t* partials, unsigned int* ticket, fr_t* out4) {
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ fr_t smem[EvalAcc<2>::NV * 32];
    __shared__ uint64_t bars[PF_WARPS];
    WarpPrefetch<4> pf; pf.init(dyn, bars);
    EvalAcc<2> acc; acc.clear();
    const size_t tiles = half / 32, nw = (size_t)gridDim.x * PF_WARPS;
    size_t tile = (size_t)blockIdx.x * PF_WARPS + (threadIdx.x >> 5);
    auto post = [&](size_t t) {
        const fr_t* rows[4] = {tabs.t[0] + t * 32, tabs.t[0] + half + t * 32, tabs.t[1] + t * 32, tabs.t[1] + half + t * 32};
        pf.issue(rows);
    };
    if (tile < tiles) post(tile);
    for (; tile < tiles; tile += nw) {
        fr_t e[4];
        pf.fetch(e);
        if (tile + nw < tiles) post(tile + nw);
        fr_t lo[2] = {e[0], e[2]}, hi[2] = {e[1], e[3]};
        acc.pair(lo, hi);
    }
    fr_t v[EvalAcc<2>::NV];
    acc.finish(v);
    grid_finish_sum<fr_t, EvalAcc<2>::NV>(v, partials, ticket, smem, EvalEpilogue<2>{out4});
}
__global__ void __launch_bounds__(SC_THREADS, 2) k_round_eval2_claim_pf(ScTables tabs, size_t half, const fr_t claim, fr_t* partials, unsigned int* ticket, fr_t* out4) {
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ fr_t smem[2 * 32];
    __shared__ uint64_t bars[PF_WARPS];
    WarpPrefetch<4> pf; pf.init(dyn, bars);
    EvalAcc2Claim acc; acc.clear();
    const size_t tiles = half / 32, nw = (size_t)gridDim.x * PF_WARPS;
    size_t tile = (size_t)blockIdx.x * PF_WARPS + (threadIdx.x >> 5);
    auto post = [&](size_t t) {
        const fr_t* rows[4] = {tabs.t[0] + t * 32, tabs.t[0] + half + t * 32, tabs.t[1] + t * 32, tabs.t[1] + half + t * 32};
        pf.issue(rows);
    };
    if (tile < tiles) post(tile);
    for (; tile < tiles; tile += nw) {
        fr_t e[4];
        pf.fetch(e);
        if (tile + nw < tiles) post(tile + nw);
        fr_t lo[2] = {e[0], e[2]}, hi[2] = {e[1], e[3]};
        acc.pair(lo, hi);
    }
    fr_t v[2];
    acc.finish(v);
    grid_finish_sum<fr_t, 2>(v, partials, ticket, smem, EvalClaimEpilogue{out4, claim});
}
// ---------------------------------------------------------------- persistent tail (d = 2, claim form)
// One CTA, tables in shared memory.  Per round: g(0) and g(2) over the pairs (p, p + m / 2) -> block reduction -> (sharded: summed with the peers
// by warp 0, as in the large-table kernels) -> published to the host mailbox; thread 0 then polls the mailbox for the next challenge's constant
// table, everybody folds in place, and so on until one entry per table is left; those two values go to the mailbox and to A[0], B[0].
constexpr int SC_TAIL_THREADS = 512;    // 128 registers per thread: the two lazy accumulators stay in registers (122 used)
constexpr unsigned long long TAIL_HOST_TIMEOUT_CYCLES = 40000000000ull;   // ~20 s at 2 GHz: the host may be slow between two calls of a round-stepped caller, but a dead host must not pin the GPU
__global__ void __launch_bounds__(SC_TAIL_THREADS, 1) k_sc_tail2(fr_t* A, fr_t* B, size_t n, const fr_ctab r0, ScTailBox* box) {
    extern __shared__ __align__(32) unsigned char tail_smem[];
    __shared__ fr_t red[2 * 32];
    __shared__ fr_ctab rc;
    __shared__ unsigned s_state;
    size_t m = n / 2;
    fr_t* sa = (fr_t*)tail_smem; fr_t* sb = sa + m;
    for (size_t p = threadIdx.x; p < m; p += blockDim.x) {          // first fold straight from HBM
        const fr_t a0 = ld256(A + p), a1 = ld256(A + p + m), b0 = ld256(B + p), b1 = ld256(B + p + m);
        sa[p] = a0 + r0.mul(a1 - a0); sb[p] = b0 + r0.mul(b1 - b0);
    }
    __syncthreads();
    unsigned seq = 0;
    while (m >= 2) {
        const size_t h = m / 2;
        EvalAcc2Claim acc; acc.clear();
        for (size_t p = threadIdx.x; p < h; p += blockDim.x) {
            const fr_t lo[2] = {sa[p], sb[p]}, hi[2] = {sa[p + h], sb[p + h]};
            acc.pair(lo, hi);
        }
        fr_t v[2];
        acc.finish(v);
        block_reduce_sum<fr_t, 2>(v, red);
        if (threadIdx.x < 32) {
            if (g_peer.on) peer_sum<2>(v, red);
            if (threadIdx.x == 0) {
                ++seq;
                volatile unsigned long long* o = box->out;
#pragma unroll
                for (int k = 0; k < 2; ++k)
#pragma unroll
                    for (int i = 0; i < 4; ++i) o[4 * k + i] = (unsigned long long)v[k].l[2 * i] | ((unsigned long long)v[k].l[2 * i + 1] << 32);
                __threadfence_system();
                box->out_seq = seq;
                const long long t0 = clock64();
                unsigned st;
                while ((st = box->in_seq) != seq && st < SC_TAIL_FLUSH) {
                    if ((unsigned long long)(clock64() - t0) > TAIL_HOST_TIMEOUT_CYCLES) { box->err = 1; st = SC_TAIL_ABORT; break; }
                }
                __threadfence_system();
                s_state = st;
            }
        }
        __syncthreads();
        if (s_state == SC_TAIL_ABORT) return;
        if (s_state == SC_TAIL_FLUSH) {                              // hand the tables back as they stand (m entries each)
            for (size_t p = threadIdx.x; p < m; p += blockDim.x) { st256(A + p, sa[p]); st256(B + p, sb[p]); }
            return;
        }
        if (threadIdx.x < 64) ((uint32_t*)rc.t)[threadIdx.x] = ((volatile uint32_t*)box->ctab)[threadIdx.x];
        __syncthreads();
        for (size_t p = threadIdx.x; p < h; p += blockDim.x) {
            const fr_t a0 = sa[p], a1 = sa[p + h], b0 = sb[p], b1 = sb[p + h];
            sa[p] = a0 + rc.mul(a1 - a0); sb[p] = b0 + rc.mul(b1 - b0);
        }
        __syncthreads();
        m = h;
    }
    if (threadIdx.x == 0) {
        st256(A, sa[0]); st256(B, sb[0]);
        volatile unsigned long long* o = box->out;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[i] = (unsigned long long)sa[0].l[2 * i] | ((unsigned long long)sa[0].l[2 * i + 1] << 32);
            o[4 + i] = (unsigned long long)sb[0].l[2 * i] | ((unsigned long long)sb[0].l[2 * i + 1] << 32);
        }
        __threadfence_system();
        box->out_seq = seq + 1;
    }
}

// positions per launch from which the prefetch variants run (tuning "prefetch_min_log2").  Measured on B200 (tools/bench_fold.py): +5.7% on the fused
// kernel and +3% on the evaluation kernel at 2^26 entries, +3-5% down to 2^22, neutral below (the tables then sit in the 126 MB L2).
static size_t g_pf_min_work = (size_t)1 << 21;
void set_prefetch_min_work(size_t w) { g_pf_min_work = w < 1024 ? 1024 : w; }

template <class K>
static cudaError_t enable_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// fr_ctab::make with native 64-bit host arithmetic (the generic one runs ~19 products of the 32-bit limb code with an emulated carry flag,
// ~5 us per round): T[k] = canonical(r * 2^(32 k + 64)), the same limbs
static fr_ctab make_ctab_host(const fr_t& r_elem) {
    using host::Fr64;
    static const Fr64 w = Fr64::from_u64(1ull << 32);
    fr_ctab c;
    Fr64 cur = Fr64::from_raw(r_elem.l) * w * w;
    for (int k = 0; k < 8; ++k) {
        const Fr64 can = cur.from_mont();
        memcpy(c.t[k], can.l, 32);
        cur = cur * w;
    }
    return c;
}

// ---------------------------------------------------------------- persistent tail: host side
cudaError_t launch_sc_tail(fr_t* A, fr_t* B, size_t n, const fr_t& r_elem, ScTailBox* box, cudaStream_t s) {
    if (n < 4 || (n / 2) > ((size_t)1 << SC_TAIL_MAX_LOG)) return cudaErrorInvalidValue;
    const size_t sm = n * sizeof(fr_t);                     // two tables of n / 2 entries
    cudaError_t e = enable_smem(k_sc_tail2, sm);
    if (e) return e;
    box->out_seq = 0; box->in_seq = 0; box->err = 0;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    k_sc_tail2<<<1, SC_TAIL_THREADS, sm, s>>>(A, B, n, make_ctab_host(r_elem), box);
    return cudaGetLastError();
}
void sc_tail_post_challenge(ScTailBox* box, const fr_t& r_elem, unsigned seq) {
    const fr_ctab c = make_ctab_host(r_elem);
    memcpy(box->ctab, c.t, sizeof(c.t));
    std::atomic_thread_fence(std::memory_order_seq_cst);    // the table is in memory before the sequence number the kernel polls
    box->in_seq = seq;
}
void sc_tail_post_command(ScTailBox* box, unsigned command) {
    std::atomic_thread_fence(std::memory_order_seq_cst);
    box->in_seq = command;
}
bool sc_tail_wait(ScTailBox* box, unsigned seq, cudaStream_t s, fr_t out[2]) {
    const auto t0 = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while (box->out_seq != seq) {
        if (box->err) return false;
        if ((++spins & 0x3fff) == 0) {
            if (cudaStreamQuery(s) != cudaErrorNotReady) { if (box->out_seq == seq) break; return false; }   // the kernel is gone without publishing: it failed
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(10)) return false;
        }
    }
    std::atomic_thread_fence(std::memory_order_seq_cst);
    memcpy(out, (const void*)box->out, 2 * sizeof(fr_t));
    return true;
}

// ---------------------------------------------------------------- launch helpers
static inline int sc_grid(size_t work, int sm_count, int blocks_per_sm) {
    size_t need = (work + SC_THREADS - 1) / SC_THREADS;
    size_t cap = (size_t)sm_count * blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

cudaError_t launch_round_eval(int d, const ScTables& tabs, size_t n, fr_t* partials, unsigned int* ticket, fr_t* out4,
                              int sm_count, cudaStream_t s, const fr_t* claim) {
    size_t half = n / 2;
    if (claim && d == 2) {
        if (half >= g_pf_min_work && half % 32 == 0) {
            const size_t sm = (size_t)PF_WARPS * 4 * 1024;
            cudaError_t e = enable_smem(k_round_eval2_claim_pf, sm);
            if (e) return e;
            k_round_eval2_claim_pf<<<sc_grid(half, sm_count, 2), SC_THREADS, sm, s>>>(tabs, half, *claim, partials, ticket, out4);
        } else {
            k_round_eval2_claim<<<sc_grid(half, sm_count, SC_BLOCKS_PER_SM), SC_THREADS, 0, s>>>(tabs, half, *claim, partials, ticket, out4);
        }
        return cudaGetLastError();
    }
    if (d == 2 && half >= g_pf_min_work && half % 32 == 0) {
        const size_t sm = (size_t)PF_WARPS * 4 * 1024;
        cudaError_t e = enable_smem(k_round_eval2_pf, sm);
        if (e) return e;
        k_round_eval2_pf<<<sc_grid(half, sm_count, 2), SC_THREADS, sm, s>>>(tabs, half, partials, ticket, out4);
        return cudaGetLastError();
    }
    int grid = sc_grid(half, sm_count, SC_BLOCKS_PER_SM);
    switch (d) {
        case 1: k_round_eval<1><<<grid, SC_THREADS, 0, s>>>(tabs, half, partials, ticket, out4); break;
        case 2: k_round_eval<2><<<grid, SC_THREADS, 0, s>>>(tabs, half, partials, ticket, out4); break;
        case 3: k_round_eval<3><<<grid, SC_THREADS, 0, s>>>(tabs, half, partials, ticket, out4); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_bind(fr_t* t, size_t n, const fr_t& r_elem, int sm_count, cudaStream_t s) {
    const fr_ctab r = make_ctab_host(r_elem);   // T[k] = r 2^(32k+64) mod p: the fold multiplies by this one constant
    size_t half = n / 2;
    int grid = sc_grid(half, sm_count, SC_BLOCKS_PER_SM_BIND);
    k_bind<<<grid, SC_THREADS, 0, s>>>(t, half, r);
    return cudaGetLastError();
}

cudaError_t launch_bind_to(const fr_t* t, fr_t* out, size_t n, const fr_t& r_elem, int sm_count, cudaStream_t s) {
    const fr_ctab r = make_ctab_host(r_elem);
    size_t half = n / 2;
    int grid = sc_grid(half, sm_count, SC_BLOCKS_PER_SM_BIND);
    k_bind_to<<<grid, SC_THREADS, 0, s>>>(t, out, half, r);
    return cudaGetLastError();
}

cudaError_t launch_bind_eval(int d, const ScTables& tabs, size_t n, const fr_t& r_elem, const fr_t* claim, fr_t* partials, unsigned int* ticket,
                             fr_t* out4, int sm_count, cudaStream_t s) {
    const fr_ctab r = make_ctab_host(r_elem);
    size_t quarter = n / 4;
    if (claim && d == 2) {
        if (quarter >= g_pf_min_work && quarter % 32 == 0) {
            const size_t sm = (size_t)PF_WARPS * 8 * 1024;
            cudaError_t e = enable_smem(k_bind_eval2_claim_pf, sm);
            if (e) return e;
            k_bind_eval2_claim_pf<<<sc_grid(quarter, sm_count, 2), SC_THREADS, sm, s>>>(tabs, quarter, r, *claim, partials, ticket, out4);
            return cudaGetLastError();
        }
        k_bind_eval2_claim<<<sc_grid(quarter, sm_count, SC_BLOCKS_PER_SM), SC_THREADS, 0, s>>>(tabs, quarter, r, *claim, partials, ticket, out4);
        return cudaGetLastError();
    }
    int grid = sc_grid(quarter, sm_count, SC_BLOCKS_PER_SM);
    switch (d) {
        case 1: k_bind_eval<1><<<grid, SC_THREADS, 0, s>>>(tabs, quarter, r, partials, ticket, out4); break;
        case 2: k_bind_eval<2><<<grid, SC_THREADS, 0, s>>>(tabs, quarter, r, partials, ticket, out4); break;
        case 3: k_bind_eval<3><<<grid, SC_THREADS, 0, s>>>(tabs, quarter, r, partials, ticket, out4); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace tsg
