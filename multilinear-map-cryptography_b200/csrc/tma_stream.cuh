// tma_stream.cuh - multi-stream HBM -> shared-memory pipeline built on the TMA bulk-copy engine.
//
// The sum-check / MLE kernels are streaming passes whose per-element arithmetic (a 256-bit Montgomery
// product per 96 bytes) is close to the integer-pipe/HBM balance point, so memory latency must be fully
// decoupled from the arithmetic.  Each CTA walks its tiles through a STAGES-deep ring of shared-memory
// buffers; one elected thread posts `cp.async.bulk.shared::cluster.global` copies (SASS: UBLKCP) of the
// next tile of every input stream and the copies signal an mbarrier with their byte count.  All
// threads wait on the barrier, pull their own 32-byte element of each stream into registers, release
// the stage (`__syncthreads`) so it can be refilled immediately, and only then do the arithmetic - so
// STAGES tiles per CTA are always in flight regardless of how long the math takes.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "fp.cuh"

namespace tsg {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

// Ring of STAGES buffers, each holding NS streams x THREADS elements of 32 bytes.
// Tile t covers positions [t * THREADS, (t + 1) * THREADS); stream s of tile t starts at
// src[s] + t * THREADS.  `ntiles` must be exact (work is a multiple of THREADS).
template <int NS, int THREADS, int STAGES>
struct Pipeline {
    static constexpr uint32_t STREAM_BYTES = THREADS * 32u;
    static constexpr uint32_t STAGE_BYTES = NS * STREAM_BYTES;
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + STAGES * sizeof(uint64_t) + 128;

    unsigned char* buf;     // STAGES * STAGE_BYTES, 128-byte aligned
    uint64_t* bars;         // STAGES mbarriers
    const fr_t* src[NS];
    size_t first_tile, tile_stride, my_tiles;

    __device__ __forceinline__ void init(unsigned char* smem_raw, const fr_t* const (&streams)[NS], size_t ntiles) {
        uintptr_t a = ((uintptr_t)smem_raw + 127) & ~(uintptr_t)127;
        buf = (unsigned char*)a;
        bars = (uint64_t*)(buf + (size_t)STAGES * STAGE_BYTES);
#pragma unroll
        for (int s = 0; s < NS; ++s) src[s] = streams[s];
        first_tile = blockIdx.x; tile_stride = gridDim.x;
        my_tiles = ntiles > first_tile ? (ntiles - first_tile + tile_stride - 1) / tile_stride : 0;
        if (threadIdx.x == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
            fence_barrier_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 0; k < STAGES && (size_t)k < my_tiles; ++k) issue(k);
        }
    }
    // tile index (global) of this CTA's k-th tile
    __device__ __forceinline__ size_t tile_of(size_t k) const { return first_tile + k * tile_stride; }

    __device__ __forceinline__ void issue(size_t k) {
        const int stage = (int)(k % STAGES);
        const size_t base = tile_of(k) * THREADS;
        mbar_expect_tx(&bars[stage], STAGE_BYTES);
#pragma unroll
        for (int s = 0; s < NS; ++s)
            bulk_g2s(buf + (size_t)stage * STAGE_BYTES + (size_t)s * STREAM_BYTES, src[s] + base, STREAM_BYTES, &bars[stage]);
    }
    // wait for tile k, copy this thread's element of every stream into registers, release the stage and
    // refill it with tile k + STAGES
    __device__ __forceinline__ void fetch(size_t k, fr_t (&e)[NS]) {
        const int stage = (int)(k % STAGES);
        mbar_wait(&bars[stage], (uint32_t)((k / STAGES) & 1));
        const fr_t* tile = (const fr_t*)(buf + (size_t)stage * STAGE_BYTES);
#pragma unroll
        for (int s = 0; s < NS; ++s) e[s] = tile[s * THREADS + threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0 && k + STAGES < my_tiles) issue(k + STAGES);
    }
};

}  // namespace tma
}  // namespace tsg
