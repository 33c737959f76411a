// fr_device.cuh - device-side helpers shared by the Fr kernels: 256-bit global loads/stores of one
// field element (LDG.E.256 / STG.E.256 on sm_100a), warp/block reductions of field elements, and the
// "last block finishes" grid reduction used by every kernel that ends in a handful of Fr sums.
#pragma once
#include <cuda_runtime.h>
#include <type_traits>
#include <utility>
#include "fp.cuh"

namespace tsg {

// one 32-byte element per instruction; p must be 32-byte aligned
template <class F>
__device__ __forceinline__ F ld256(const F* p) {
    F r;
    asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
                 : "l"(p));
    return r;
}
// read-only data that no thread of the running kernel writes
template <class F>
__device__ __forceinline__ F ld256_nc(const F* p) {
    F r;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
                 : "l"(p));
    return r;
}
// streaming variant: data touched once, do not keep in L1
template <class F>
__device__ __forceinline__ F ld256_stream(const F* p) {
    F r;
    asm volatile("ld.global.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
                 : "l"(p));
    return r;
}
// L2-coherent load (bypasses L1) for data produced by other blocks of the running kernel
template <class F>
__device__ __forceinline__ F ld256_cg(const F* p) {
    F r;
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
                 : "l"(p) : "memory");
    return r;
}
template <class F>
__device__ __forceinline__ void st256(F* p, const F& v) {
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(v.l[0]), "r"(v.l[1]), "r"(v.l[2]), "r"(v.l[3]), "r"(v.l[4]), "r"(v.l[5]), "r"(v.l[6]), "r"(v.l[7])
                 : "memory");
}

template <class F>
__device__ __forceinline__ F shfl_down_fp(const F& v, int delta) {
    F r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = __shfl_down_sync(0xffffffffu, v.l[i], delta);
    return r;
}

// Sum NV field elements over the block (modular adds).  Result valid in thread 0.  smem: NV * 32 elements.
template <class F, int NV>
__device__ __forceinline__ void block_reduce_sum(F (&v)[NV], F* smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v[k] = v[k] + shfl_down_fp(v[k], d);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) smem[warp * NV + k] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            F x = lane < nwarps ? smem[lane * NV + k] : F::zero();
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) x = x + shfl_down_fp(x, d);
            v[k] = x;
        }
    }
}

// Grid-wide finish: every block deposits NV partial sums; the last block to arrive (atomic ticket)
// adds them all and writes `out[0..NV)`.  `partials` holds gridDim.x * NV elements, `ticket` is a
// zero-initialised counter that the finishing block resets for the next launch.
// An epilogue with a member `warp(F (&v)[NV], F* smem)` is run by the 32 lanes of warp 0 of the finishing block (v valid in lane 0, smem free for its
// use) instead of by thread 0 alone - the sharded sum-check uses the lanes to talk to all peers at once.
template <class E, class F, int NV, class = void> struct epilogue_has_warp : std::false_type {};
template <class E, class F, int NV>
struct epilogue_has_warp<E, F, NV, std::void_t<decltype(std::declval<const E&>().warp(std::declval<F (&)[NV]>(), (F*)nullptr))>> : std::true_type {};

template <class F, int NV, class Epilogue>
__device__ __forceinline__ void grid_finish_sum(F (&v)[NV], F* partials, unsigned int* ticket, F* smem, Epilogue epi) {
    __shared__ bool s_last;
    block_reduce_sum<F, NV>(v, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) partials[blockIdx.x * NV + k] = v[k];
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    F acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = F::zero();
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] = acc[k] + ld256_cg(partials + b * NV + k);
    }
    __syncthreads();
    block_reduce_sum<F, NV>(acc, smem);
    if constexpr (epilogue_has_warp<Epilogue, F, NV>::value) {
        if (threadIdx.x < 32) epi.warp(acc, smem);
        if (threadIdx.x == 0) *ticket = 0;
    } else if (threadIdx.x == 0) {
        epi(acc);
        *ticket = 0;
    }
}

}  // namespace tsg
