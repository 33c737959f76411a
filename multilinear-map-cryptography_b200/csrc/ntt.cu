// ntt.cu - batched radix-2 number-theoretic transforms over BN254 Fr (two-adicity 28, generator 5).
//
// Used by the fast interpolation on {0..n-1} (interp.cu) that replaces the reference's O(n^3)
// poly_utils::lagrange_interpolate (src/polynomials.rs:301-352) on the prove path.
//   forward: Gentleman-Sande DIF, natural order in -> bit-reversed out
//   inverse: Cooley-Tukey DIT, bit-reversed in -> natural out, NOT scaled by 1/m (callers fold the scale
//            into the spectrum they multiply with)
// `data` holds `batch` contiguous arrays of m = 2^logm elements.  Stages whose butterfly span fits a
// 1024-element chunk run in shared memory in one kernel; wider stages are one streaming pass each.
#include "fr_device.cuh"
#include "ntt.cuh"

namespace tsg {

// one DIF stage with half-span 2^h over all arrays: (u, v) -> (u + v, (u - v) w^j)
__global__ void __launch_bounds__(256) k_ntt_dif_stage(fr_t* data, unsigned logm, unsigned h, const fr_t* tw, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t half = (size_t)1 << h, mh = (size_t)1 << (logm - 1);
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        size_t a = t >> (logm - 1), u = t & (mh - 1);
        size_t blk = u >> h, j = u & (half - 1);
        size_t i0 = (a << logm) + (blk << (h + 1)) + j, i1 = i0 + half;
        fr_t x = ld256(data + i0), y = ld256(data + i1);
        fr_t w = ld256_nc(tw + (j << (logm - 1 - h)));
        st256(data + i0, x + y);
        st256(data + i1, (x - y) * w);
    }
}
// one DIT stage with half-span 2^h: (u, v) -> (u + v w^j, u - v w^j)
__global__ void __launch_bounds__(256) k_ntt_dit_stage(fr_t* data, unsigned logm, unsigned h, const fr_t* twi, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t half = (size_t)1 << h, mh = (size_t)1 << (logm - 1);
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        size_t a = t >> (logm - 1), u = t & (mh - 1);
        size_t blk = u >> h, j = u & (half - 1);
        size_t i0 = (a << logm) + (blk << (h + 1)) + j, i1 = i0 + half;
        fr_t x = ld256(data + i0);
        fr_t y = ld256(data + i1) * ld256_nc(twi + (j << (logm - 1 - h)));
        st256(data + i0, x + y);
        st256(data + i1, x - y);
    }
}

// last `nst` DIF stages (spans 2^nst .. 2) on chunks of NTT_CHUNK contiguous elements in shared memory
__global__ void __launch_bounds__(NTT_THREADS) k_ntt_dif_tail(fr_t* data, unsigned logm, unsigned nst, const fr_t* tw, size_t total_elems) {
    extern __shared__ __align__(16) unsigned char ntt_smem[];
    fr_t* sh = (fr_t*)ntt_smem;
    const size_t base = (size_t)blockIdx.x * NTT_CHUNK;
    const size_t cnt = total_elems - base < NTT_CHUNK ? total_elems - base : NTT_CHUNK;
    for (size_t i = threadIdx.x; i < cnt; i += blockDim.x) sh[i] = ld256(data + base + i);
    __syncthreads();
    for (unsigned l = nst; l >= 1; --l) {
        const size_t half = (size_t)1 << (l - 1);
        for (size_t u = threadIdx.x; u < cnt / 2; u += blockDim.x) {
            size_t blk = u >> (l - 1), j = u & (half - 1);
            size_t i0 = (blk << l) + j, i1 = i0 + half;
            fr_t x = sh[i0], y = sh[i1];
            fr_t w = ld256_nc(tw + (j << (logm - l)));
            sh[i0] = x + y;
            sh[i1] = (x - y) * w;
        }
        __syncthreads();
    }
    for (size_t i = threadIdx.x; i < cnt; i += blockDim.x) st256(data + base + i, sh[i]);
}
// first `nst` DIT stages (spans 2 .. 2^nst)
__global__ void __launch_bounds__(NTT_THREADS) k_ntt_dit_head(fr_t* data, unsigned logm, unsigned nst, const fr_t* twi, size_t total_elems) {
    extern __shared__ __align__(16) unsigned char ntt_smem[];
    fr_t* sh = (fr_t*)ntt_smem;
    const size_t base = (size_t)blockIdx.x * NTT_CHUNK;
    const size_t cnt = total_elems - base < NTT_CHUNK ? total_elems - base : NTT_CHUNK;
    for (size_t i = threadIdx.x; i < cnt; i += blockDim.x) sh[i] = ld256(data + base + i);
    __syncthreads();
    for (unsigned l = 1; l <= nst; ++l) {
        const size_t half = (size_t)1 << (l - 1);
        for (size_t u = threadIdx.x; u < cnt / 2; u += blockDim.x) {
            size_t blk = u >> (l - 1), j = u & (half - 1);
            size_t i0 = (blk << l) + j, i1 = i0 + half;
            fr_t x = sh[i0];
            fr_t y = sh[i1] * ld256_nc(twi + (j << (logm - l)));
            sh[i0] = x + y;
            sh[i1] = x - y;
        }
        __syncthreads();
    }
    for (size_t i = threadIdx.x; i < cnt; i += blockDim.x) st256(data + base + i, sh[i]);
}

static inline int gridfor(size_t work, int threads, size_t cap) {
    size_t g = (work + threads - 1) / threads;
    if (g < 1) g = 1;
    return (int)(g < cap ? g : cap);
}

static bool g_smem_set = false;
static cudaError_t ensure_smem() {
    if (g_smem_set) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k_ntt_dif_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NTT_CHUNK * sizeof(fr_t)));
    if (e) return e;
    e = cudaFuncSetAttribute(k_ntt_dit_head, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NTT_CHUNK * sizeof(fr_t)));
    if (e) return e;
    g_smem_set = true;
    return cudaSuccess;
}

cudaError_t ntt_forward(fr_t* data, unsigned logm, size_t batch, const fr_t* tw, int sm_count, cudaStream_t s, unsigned* launches) {
    if (logm == 0) return cudaSuccess;
    cudaError_t e = ensure_smem();
    if (e) return e;
    const size_t total = batch << logm;
    const unsigned nst = logm < NTT_LOG_CHUNK ? logm : NTT_LOG_CHUNK;
    for (unsigned h = logm; h-- > nst;) {   // half-span 2^h, from the widest stage down
        k_ntt_dif_stage<<<gridfor(total / 2, 256, (size_t)sm_count * 8), 256, 0, s>>>(data, logm, h, tw, total / 2);
        if (launches) ++*launches;
    }
    size_t chunks = (total + NTT_CHUNK - 1) / NTT_CHUNK;
    k_ntt_dif_tail<<<(unsigned)chunks, NTT_THREADS, NTT_CHUNK * sizeof(fr_t), s>>>(data, logm, nst, tw, total);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t ntt_inverse_unscaled(fr_t* data, unsigned logm, size_t batch, const fr_t* twi, int sm_count, cudaStream_t s, unsigned* launches) {
    if (logm == 0) return cudaSuccess;
    cudaError_t e = ensure_smem();
    if (e) return e;
    const size_t total = batch << logm;
    const unsigned nst = logm < NTT_LOG_CHUNK ? logm : NTT_LOG_CHUNK;
    size_t chunks = (total + NTT_CHUNK - 1) / NTT_CHUNK;
    k_ntt_dit_head<<<(unsigned)chunks, NTT_THREADS, NTT_CHUNK * sizeof(fr_t), s>>>(data, logm, nst, twi, total);
    if (launches) ++*launches;
    for (unsigned h = nst; h < logm; ++h) {
        k_ntt_dit_stage<<<gridfor(total / 2, 256, (size_t)sm_count * 8), 256, 0, s>>>(data, logm, h, twi, total / 2);
        if (launches) ++*launches;
    }
    return cudaGetLastError();
}

}  // namespace tsg
