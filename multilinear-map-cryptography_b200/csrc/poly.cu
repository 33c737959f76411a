// poly.cu - univariate kernels of KZGCommitment::open: value P(z) and quotient (P - P(z)) / (x - z).
//
// Replaces evaluate_polynomial (Horner, src/commitments.rs:305-313) and compute_quotient_polynomial /
// polynomial_division (src/commitments.rs:317-375), two serial O(n) recurrences, by one suffix recurrence
//     H_k = c_k + z H_{k+1}   (H_n = 0):   value = H_0,  quotient q_i = H_{i+1}
// computed as a three-level blocked scan: each thread runs Horner over POLY_SPAN coefficients, a block
// combines its threads with a weighted suffix scan (weights z^(SPAN 2^s)), a single block combines the
// block totals, and the second kernel replays each thread's span with the correct incoming suffix value.
// HBM traffic: 32 n read twice + 32 n written.
#include "fr_device.cuh"
#include "poly.cuh"

namespace tsg {

// local Horner over this thread's span: C = sum_{j in span} c_j z^(j - begin)
__device__ __forceinline__ fr_t span_horner(const fr_t* c, size_t b, size_t e, const fr_t& z) {
    fr_t acc = fr_t::zero();
    for (size_t j = e; j-- > b;) acc = acc * z + ld256_nc(c + j);
    return acc;
}

// weighted suffix scan inside a block: X_t <- sum_{u >= t} X_u w^(u - t), w = z^SPAN.  pw[s] = w^(2^s).
__device__ __forceinline__ fr_t block_suffix_scan(fr_t x, const fr_t* pw, fr_t* sh) {
    const int t = threadIdx.x, nt = blockDim.x;
    sh[t] = x;
    __syncthreads();
    int s = 0;
    for (int d = 1; d < nt; d <<= 1, ++s) {
        fr_t other = (t + d < nt) ? sh[t + d] : fr_t::zero();
        __syncthreads();
        if (t + d < nt) x = x + pw[s] * other;
        sh[t] = x;
        __syncthreads();
    }
    return x;
}

// pass 1: block totals  B_blk = sum_{j in block} c_j z^(j - block_begin)
__global__ void __launch_bounds__(POLY_THREADS) k_poly_block_totals(const fr_t* c, size_t n, fr_t z, const fr_t* pw, fr_t* totals) {
    __shared__ fr_t sh[POLY_THREADS];
    const size_t blk_begin = (size_t)blockIdx.x * POLY_THREADS * POLY_SPAN;
    size_t b = blk_begin + (size_t)threadIdx.x * POLY_SPAN, e = b + POLY_SPAN;
    if (b > n) b = n;
    if (e > n) e = n;
    fr_t x = span_horner(c, b, e, z);
    x = block_suffix_scan(x, pw, sh);
    if (threadIdx.x == 0) st256(totals + blockIdx.x, x);
}

// pass 2 (single block): carry[blk] = H at the END of block blk = sum_{u > blk} totals[u] W^(u - blk - 1), W = z^(THREADS*SPAN)
__global__ void k_poly_carry(const fr_t* totals, size_t nblocks, fr_t W, fr_t* carry) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    fr_t acc = fr_t::zero();
    for (size_t u = nblocks; u-- > 0;) {
        st256(carry + u, acc);
        acc = acc * W + ld256(totals + u);
    }
}

// pass 3: replay.  q[i] = H_{i+1} for i in [0, n-1), value = H_0.
__global__ void __launch_bounds__(POLY_THREADS) k_poly_apply(const fr_t* c, size_t n, fr_t z, const fr_t* pw, const fr_t* carry, fr_t zspan,
                                                            fr_t* q, fr_t* value) {
    __shared__ fr_t sh[POLY_THREADS];
    const size_t blk_begin = (size_t)blockIdx.x * POLY_THREADS * POLY_SPAN;
    size_t b = blk_begin + (size_t)threadIdx.x * POLY_SPAN, e = b + POLY_SPAN;
    if (b > n) b = n;
    if (e > n) e = n;
    fr_t x = span_horner(c, b, e, z);
    // inclusive weighted suffix over threads, then convert to "suffix value entering my span from the right"
    fr_t incl = block_suffix_scan(x, pw, sh);
    __syncthreads();
    sh[threadIdx.x] = incl;
    __syncthreads();
    // H at the end of my span = (inclusive value of the next thread) + carry-in of the block scaled to that position
    // H_{end of block} = carry[blk]; for thread t: H_end(t) = incl[t+1] + w^(nt-1-t) * carry  where w = z^SPAN
    // computed incrementally instead: walk from the block end is serial, so use incl and powers:
    fr_t cin = ld256_nc(carry + blockIdx.x);
    // w^(nt - 1 - t): square-and-multiply over the bits of (nt-1-t) using pw[]
    fr_t wp = fr_t::one();
    unsigned ex = blockDim.x - 1 - threadIdx.x;
    for (int s = 0; ex; ++s, ex >>= 1) if (ex & 1) wp = wp * pw[s];
    fr_t h = ((threadIdx.x + 1 < blockDim.x) ? sh[threadIdx.x + 1] : fr_t::zero()) + wp * cin;
    (void)zspan;
    // replay my span from the top: H_j = c_j + z H_{j+1}
    for (size_t j = e; j-- > b;) {
        h = h * z + ld256_nc(c + j);
        if (j > 0) st256(q + j - 1, h); else st256(value, h);
    }
}

cudaError_t poly_open_launch(const fr_t* c, size_t n, const fr_t& z, const fr_t* pw_dev, const fr_t& W, fr_t* totals, fr_t* carry,
                             fr_t* q, fr_t* value, cudaStream_t s, unsigned* launches) {
    const size_t per_block = (size_t)POLY_THREADS * POLY_SPAN;
    const size_t nblocks = (n + per_block - 1) / per_block;
    k_poly_block_totals<<<(unsigned)nblocks, POLY_THREADS, 0, s>>>(c, n, z, pw_dev, totals);
    k_poly_carry<<<1, 32, 0, s>>>(totals, nblocks, W, carry);
    k_poly_apply<<<(unsigned)nblocks, POLY_THREADS, 0, s>>>(c, n, z, pw_dev, carry, W, q, value);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

}  // namespace tsg
