// lookup.cu - kernels of the lookup (read-checking) argument: the sparse one-hot matrix ra(x, j) = [idx_j == x] of Shout is never
// materialised (K x T entries); what the sum-check needs is its product with a weight vector,
//     A[x] = sum_{j : idx_j = x} W[j]                      (A = ra~(., r) for W = eq(r, .)),
// and the gather G[j] = S[idx_j] for the verifier's closing check.  This is the constraint the reference leaves as a stub
// (src/shout.rs:157-184: "In a production implementation, this would involve more complex constraints ...").
//
// Field elements cannot be added atomically, but their limbs can: the Montgomery representation is linear, so A[x] is the sum
// of the 256-bit integers W[j] reduced mod r once at the end.  Every bucket owns eight 64-bit counters, one per 32-bit limb
// (exact for up to 2^32 contributions per bucket); a second kernel carries and reduces them.  The result does not depend on the
// order of the atomics, so it is deterministic.  Tables are in the bit-reversed position order of sumcheck.cu.
#include "fr_device.cuh"
#include "lookup.cuh"

namespace tsg {

__device__ __forceinline__ unsigned long long bitrev_u64(unsigned long long x, unsigned bits) {
    return bits ? (__brevll(x) >> (64 - bits)) : 0ull;
}

// j runs in natural order: the index loads are coalesced, the weight loads are whole 32-byte sectors
__global__ void __launch_bounds__(256) k_weighted_hist(const fr_t* W, unsigned l, const unsigned long long* idx, size_t n, unsigned long long* acc) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const fr_t w = ld256_nc(W + bitrev_u64(j, l));
        unsigned long long* a = acc + 8 * idx[j];
#pragma unroll
        for (int i = 0; i < 8; ++i) if (w.l[i]) atomicAdd(a + i, (unsigned long long)w.l[i]);
    }
}

__global__ void __launch_bounds__(256) k_limb_sums_to_table(const unsigned long long* acc, unsigned k, fr_t* out) {
    const size_t K = (size_t)1 << k;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < K; x += stride) {
        // carry-propagate the eight limb sums into 256 low bits + a 64-bit high part
        fr_t lo;
        unsigned long long carry = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const unsigned long long s = acc[8 * x + i];
            const unsigned long long v = (s & 0xffffffffull) + (carry & 0xffffffffull);
            lo.l[i] = (uint32_t)v;
            carry = (s >> 32) + (carry >> 32) + (v >> 32);
        }
        // low part < 2^256 < 6 r: at most five subtractions; high part: hi * 2^256 mod r is the Montgomery form of hi
#pragma unroll
        for (int t = 0; t < 5; ++t) limb::cond_sub_mod<FrP>(lo.l);
        fr_t r = lo;
        if (carry) r = r + fr_t::from_u64(carry);
        st256(out + bitrev_u64(x, k), r);
    }
}

__global__ void __launch_bounds__(256) k_table_gather(const fr_t* src, unsigned k, const unsigned long long* idx, size_t n, unsigned l, fr_t* out) {
    const size_t L = (size_t)1 << l;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < L; j += stride) {
        fr_t v = fr_t::zero();
        if (j < n) v = ld256_nc(src + bitrev_u64(idx[j], k));
        st256(out + bitrev_u64(j, l), v);
    }
}

static inline int grid_for(size_t work, int threads, size_t cap) {
    size_t g = (work + threads - 1) / threads;
    if (g < 1) g = 1;
    return (int)(g < cap ? g : cap);
}
cudaError_t launch_weighted_hist(const fr_t* W, unsigned l, const unsigned long long* idx, size_t n, unsigned long long* acc, int sm_count, cudaStream_t s) {
    if (!n) return cudaSuccess;
    k_weighted_hist<<<grid_for(n, 256, (size_t)sm_count * 8), 256, 0, s>>>(W, l, idx, n, acc);
    return cudaGetLastError();
}
cudaError_t launch_limb_sums_to_table(const unsigned long long* acc, unsigned k, fr_t* out, int sm_count, cudaStream_t s) {
    k_limb_sums_to_table<<<grid_for((size_t)1 << k, 256, (size_t)sm_count * 8), 256, 0, s>>>(acc, k, out);
    return cudaGetLastError();
}
cudaError_t launch_table_gather(const fr_t* src, unsigned k, const unsigned long long* idx, size_t n, unsigned l, fr_t* out, int sm_count, cudaStream_t s) {
    k_table_gather<<<grid_for((size_t)1 << l, 256, (size_t)sm_count * 8), 256, 0, s>>>(src, k, idx, n, l, out);
    return cudaGetLastError();
}

}  // namespace tsg

// ================================================================ read/write memory (Twist) tables
// Dense tables over (cell x, cycle j), reference index i = x + K j (K = 2^k cells in the low variables, T = 2^t cycles in the high
// ones: the layout of BASELINE config 4), stored at the bit-reversed position (bitrev_k(x) << t) | bitrev_t(j).
namespace tsg {

// Val(x, j) = content of cell x just before operation j (zero-initialised memory; for j >= n: the final content).  A cell's row changes only
// at the writes to it: write number w (operation wj[w], address wa[w], value wv[w]) holds for the cycles wj[w] + 1 .. wnext[w] - 1, where
// wnext[w] is the next write to the same address (or T).  One block per write fills that run of the pre-zeroed table; the runs are disjoint
// and cover everything that is not zero.  (The host derives wnext in one pass over the trace.)
__global__ void __launch_bounds__(256) k_val_fill(const unsigned long long* wa, const unsigned long long* wj, const unsigned long long* wnext, const fr_t* wv,
                                                  size_t nwrites, unsigned k, unsigned t, fr_t* out) {
    for (size_t w = blockIdx.x; w < nwrites; w += gridDim.x) {
        const fr_t v = ld256_nc(wv + w);
        if (v.is_zero()) continue;
        fr_t* row = out + (bitrev_u64(wa[w], k) << t);
        const unsigned long long end = wnext[w];
        for (unsigned long long j = wj[w] + 1 + threadIdx.x; j < end; j += blockDim.x) st256(row + bitrev_u64(j, t), v);
    }
}

// out[(x = addr[j]) + K j] = W[j] for every j < n with keep[j] == flag (table pre-zeroed): the one-hot matrix with weighted rows
__global__ void __launch_bounds__(256) k_one_hot_weighted(const fr_t* W, const unsigned long long* addr, const unsigned char* sel, unsigned char flag, size_t n,
                                                          unsigned k, unsigned t, fr_t* out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        if (sel[j] != flag) continue;
        const unsigned long long jr = bitrev_u64(j, t);
        st256(out + ((bitrev_u64(addr[j], k) << t) | jr), ld256_nc(W + jr));
    }
}

// inout[(x, j)] = rows[j] - inout[(x, j)]: a per-cycle vector broadcast over the cells minus the table (write-checking: wv(j) - Val(x, j)).
// Position (bitrev_k(x) << t) | bitrev_t(j): the row value of a position is rows[pos & (2^t - 1)] - contiguous, coalesced.
__global__ void __launch_bounds__(256) k_broadcast_rows_minus(const fr_t* rows, unsigned t, fr_t* inout, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t mask = ((size_t)1 << t) - 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) st256(inout + i, ld256_nc(rows + (i & mask)) - ld256_stream(inout + i));
}

// out = a * b elementwise (full Montgomery products; layout-agnostic)
__global__ void __launch_bounds__(256) k_table_mul(const fr_t* a, const fr_t* b, fr_t* out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) st256(out + i, ld256_stream(a + i) * ld256_stream(b + i));
}

// V[a] = LT~(a, b) for every boolean a in {0,1}^t and a field point b (device array of t elements): the multilinear extension in the
// second argument of [a < c] (integers, bit t - 1 most significant), evaluated at c = b:
//     LT~(a, b) = sum_i (1 - a_i) b_i prod_{l > i} eq(a_l, b_l),   eq(a_l, b_l) = a_l ? b_l : 1 - b_l
__global__ void __launch_bounds__(256) k_lt_point_table(const fr_t* b, unsigned t, fr_t* out) {
    const size_t T = (size_t)1 << t;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t pos = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pos < T; pos += stride) {
        const unsigned long long a = bitrev_u64(pos, t);
        fr_t prefix = fr_t::one(), acc = fr_t::zero();
        for (int i = (int)t - 1; i >= 0; --i) {
            const fr_t bi = b[i];
            if ((a >> i) & 1) prefix = prefix * bi;
            else { acc = acc + prefix * bi; prefix = prefix * (fr_t::one() - bi); }
        }
        st256(out + pos, acc);
    }
}

cudaError_t launch_val_fill(const unsigned long long* wa, const unsigned long long* wj, const unsigned long long* wnext, const fr_t* wv, size_t nwrites,
                            unsigned k, unsigned t, fr_t* out, int sm_count, cudaStream_t s) {
    if (!nwrites) return cudaSuccess;
    size_t cap = (size_t)sm_count * 16;
    k_val_fill<<<(unsigned)(nwrites < cap ? nwrites : cap), 256, 0, s>>>(wa, wj, wnext, wv, nwrites, k, t, out);
    return cudaGetLastError();
}
cudaError_t launch_one_hot_weighted(const fr_t* W, const unsigned long long* addr, const unsigned char* sel, unsigned char flag, size_t n, unsigned k, unsigned t,
                                    fr_t* out, int sm_count, cudaStream_t s) {
    if (!n) return cudaSuccess;
    k_one_hot_weighted<<<grid_for(n, 256, (size_t)sm_count * 8), 256, 0, s>>>(W, addr, sel, flag, n, k, t, out);
    return cudaGetLastError();
}
cudaError_t launch_broadcast_rows_minus(const fr_t* rows, unsigned t, fr_t* inout, size_t n, int sm_count, cudaStream_t s) {
    k_broadcast_rows_minus<<<grid_for(n, 256, (size_t)sm_count * 8), 256, 0, s>>>(rows, t, inout, n);
    return cudaGetLastError();
}
cudaError_t launch_table_mul(const fr_t* a, const fr_t* b, fr_t* out, size_t n, int sm_count, cudaStream_t s) {
    k_table_mul<<<grid_for(n, 256, (size_t)sm_count * 8), 256, 0, s>>>(a, b, out, n);
    return cudaGetLastError();
}
cudaError_t launch_lt_point_table(const fr_t* b_dev, unsigned t, fr_t* out, int sm_count, cudaStream_t s) {
    k_lt_point_table<<<grid_for((size_t)1 << t, 256, (size_t)sm_count * 8), 256, 0, s>>>(b_dev, t, out);
    return cudaGetLastError();
}

}  // namespace tsg
