// mle.cu - multilinear-extension kernels: evaluate, partial evaluate, table generators, layout permutation.
//
// Replaces MultilinearExtension::evaluate (src/polynomials.rs:85-122, O(n 2^n) basis products) and
// MultilinearExtension::partial_evaluate (src/polynomials.rs:126-161, 2^(n-k) full evaluations) by one
// streaming pass over the table:
//     partial_evaluate(r_0..r_{k-1})[j] = sum_{a < 2^k} eq(r, a) * T[a + 2^k j]
// In the bit-reversed HBM layout (see sumcheck.cu) the fixed variables are the TOP k position bits, so
// this is a weighted column sum over 2^k contiguous rows of 2^(n-k) entries: thread c streams column c
// of every row with 256-bit loads and accumulates eq-weight x entry products lazily in 512 bits.
// evaluate(r) is the same pass with k = n - l followed by a 2^l-entry dot product with eq over the
// remaining variables.  Algorithmic bytes: 32 N (1 + 2^-k)  (SURVEY 8d).
#include "fr_device.cuh"
#include "mle.cuh"

namespace tsg {

__device__ __forceinline__ unsigned long long bitrev64(unsigned long long x, unsigned bits) {
    return bits ? (__brevll(x) >> (64 - bits)) : 0ull;
}

// out[pos] = prod_{j<b} (bit_{b-1-j}(pos) ? w_j : 1 - w_j): eq(w, .) in bit-reversed position order
// (variable 0 is the top position bit).  Each thread builds the product over the low LOWB bits once per
// 2^LOWB-entry run?  Kept simple: one thread per entry, b multiplications (b <= 32).
__global__ void k_eq_table_bitrev(const fr_t* w, unsigned b, fr_t* out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t pos = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pos < n; pos += stride) {
        fr_t acc = fr_t::one();
        for (unsigned j = 0; j < b; ++j) {
            fr_t wj = w[j];
            bool bit = (pos >> (b - 1 - j)) & 1;
            acc = acc * (bit ? wj : fr_t::one() - wj);
        }
        st256(out + pos, acc);
    }
}

// dst[bitrev(i)] = src[i]  (also its own inverse)
__global__ void k_bitrev_permute(const fr_t* src, fr_t* dst, unsigned bits, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        st256(dst + bitrev64(i, bits), ld256_nc(src + i));
    }
}

// one-hot rows: table[bitrev(row * K + idx[row])] = 1 (Montgomery one); table pre-zeroed
__global__ void k_one_hot_scatter(const unsigned long long* idx, size_t rows, unsigned logK, unsigned bits, fr_t* table) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
        unsigned long long i = (r << logK) | (idx[r] & ((1ull << logK) - 1));
        st256(table + bitrev64(i, bits), fr_t::one());
    }
}

// u64 -> Fr (Montgomery), optionally bit-reversed destination
__global__ void k_fr_from_u64(const unsigned long long* src, size_t n, fr_t* dst, unsigned bits, int bitrev) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        fr_t v = fr_t::zero();
        unsigned long long x = src[i];
        v.l[0] = (uint32_t)x; v.l[1] = (uint32_t)(x >> 32);
        v = v.to_mont();
        st256(dst + (bitrev ? bitrev64(i, bits) : i), v);
    }
}

// sparse entries: table[bitrev(idx[i])] = vals[i]; table pre-zeroed, indices unique (the host keeps the LAST occurrence of a
// repeated index, which is what the sequential loop of MultilinearExtension::from_sparse leaves, src/polynomials.rs:52-67)
__global__ void k_sparse_scatter(const unsigned long long* idx, const fr_t* vals, size_t count, unsigned bits, fr_t* table) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        st256(table + bitrev64(idx[i], bits), ld256_nc(vals + i));
}

// LessThanPolynomial::to_multilinear_extension (src/polynomials.rs:243-263): 2n variables, reference index i = a | (b << n),
// entry = lt(a, b) where the FIRST differing bit counted from bit 0 decides (src/polynomials.rs:222-239): a < b iff at the
// lowest set bit of a ^ b the bit of b is set.  One thread per position (coalesced 256-bit stores), i = bitrev(position).
__global__ void k_lt_table(unsigned n, fr_t* out, size_t size) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const unsigned long long mask = n >= 64 ? ~0ull : ((1ull << n) - 1);
    const fr_t one = fr_t::one(), zero = fr_t::zero();
    for (size_t pos = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pos < size; pos += stride) {
        const unsigned long long i = bitrev64(pos, 2 * n);
        const unsigned long long a = i & mask, b = i >> n;
        const unsigned long long x = a ^ b, low = x & (0ull - x);
        st256(out + pos, (b & low) ? one : zero);
    }
}

// MultilinearExtension::add (src/polynomials.rs:164-176) and scalar_mul (:179-189): elementwise, so the layout does not matter.
// HBM-bound: 96 resp. 64 bytes per entry; the scalar goes through the per-launch constant table like a fold challenge.
__global__ void __launch_bounds__(MLE_THREADS) k_table_add(const fr_t* a, const fr_t* b, fr_t* out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        st256(out + i, ld256_stream(a + i) + ld256_stream(b + i));
}
__global__ void __launch_bounds__(MLE_THREADS) k_table_scale(const fr_t* a, fr_t* out, size_t n, const fr_ctab s) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        st256(out + i, s.mul(ld256_stream(a + i)));
}
// MultilinearExtension::sum_evaluations (src/polynomials.rs:192-195): one streaming pass, 32 bytes per entry
struct SumEpilogue {
    fr_t* out;
    __device__ void operator()(fr_t (&v)[1]) const { *out = v[0]; }
};
__global__ void __launch_bounds__(MLE_THREADS) k_table_sum(const fr_t* a, size_t n, fr_t* partials, unsigned int* ticket, fr_t* out) {
    __shared__ fr_t smem[32];
    fr_t acc0 = fr_t::zero(), acc1 = fr_t::zero();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n; i += 2 * stride) {         // two loads in flight, two independent add chains
        fr_t x = ld256_stream(a + i), y = ld256_stream(a + i + stride);
        acc0 = acc0 + x; acc1 = acc1 + y;
    }
    if (i < n) acc0 = acc0 + ld256_stream(a + i);
    fr_t v[1]; v[0] = acc0 + acc1;
    grid_finish_sum<fr_t, 1>(v, partials, ticket, smem, SumEpilogue{out});
}

// ---------------------------------------------------------------- weighted column sum
// out_partial[split][c] = sum_{row in split} W[row] * T[row * cols + c]
// grid.x covers columns, grid.y = row splits.
__global__ void __launch_bounds__(MLE_THREADS) k_colsum(const fr_t* T, const fr_t* W, size_t rows, size_t cols, fr_t* out_partial) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const size_t nsplit = gridDim.y, split = blockIdx.y;
    const size_t r0 = rows * split / nsplit, r1 = rows * (split + 1) / nsplit;
    wide_acc<FrP> acc; acc.clear();
    size_t r = r0;
    // 4 rows in flight per thread
    for (; r + 4 <= r1; r += 4) {
        fr_t t0 = ld256_stream(T + (r + 0) * cols + c), t1 = ld256_stream(T + (r + 1) * cols + c);
        fr_t t2 = ld256_stream(T + (r + 2) * cols + c), t3 = ld256_stream(T + (r + 3) * cols + c);
        acc.add_product(t0, ld256_nc(W + r + 0));
        acc.add_product(t1, ld256_nc(W + r + 1));
        acc.add_product(t2, ld256_nc(W + r + 2));
        acc.add_product(t3, ld256_nc(W + r + 3));
    }
    for (; r < r1; ++r) acc.add_product(ld256_stream(T + r * cols + c), ld256_nc(W + r));
    st256(out_partial + split * cols + c, acc.reduce());
}

// out[c] = sum_split partial[split][c]
__global__ void k_colsum_finish(const fr_t* partial, size_t nsplit, size_t cols, fr_t* out) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    fr_t acc = ld256_nc(partial + c);
    for (size_t s = 1; s < nsplit; ++s) acc = acc + ld256_nc(partial + s * cols + c);
    st256(out + c, acc);
}

// out[0] = sum_i A[i] * B[i]
struct DotEpilogue {
    fr_t* out;
    __device__ void operator()(fr_t (&v)[1]) const { *out = v[0]; }
};
__global__ void __launch_bounds__(MLE_THREADS) k_dot(const fr_t* A, const fr_t* B, size_t n, fr_t* partials, unsigned int* ticket, fr_t* out) {
    __shared__ fr_t smem[32];
    wide_acc<FrP> acc; acc.clear();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        acc.add_product(ld256_nc(A + i), ld256_nc(B + i));
    fr_t v[1]; v[0] = acc.reduce();
    grid_finish_sum<fr_t, 1>(v, partials, ticket, smem, DotEpilogue{out});
}

// ---------------------------------------------------------------- launchers
static inline int grid_for(size_t work, int threads, size_t cap) {
    size_t g = (work + threads - 1) / threads;
    if (g < 1) g = 1;
    return (int)(g < cap ? g : cap);
}

cudaError_t launch_eq_table_bitrev(const fr_t* w_dev, unsigned b, fr_t* out, int sm_count, cudaStream_t s) {
    size_t n = (size_t)1 << b;
    k_eq_table_bitrev<<<grid_for(n, 256, (size_t)sm_count * 8), 256, 0, s>>>(w_dev, b, out, n);
    return cudaGetLastError();
}
cudaError_t launch_bitrev_permute(const fr_t* src, fr_t* dst, unsigned bits, int sm_count, cudaStream_t s) {
    size_t n = (size_t)1 << bits;
    k_bitrev_permute<<<grid_for(n, 256, (size_t)sm_count * 8), 256, 0, s>>>(src, dst, bits, n);
    return cudaGetLastError();
}
cudaError_t launch_one_hot_scatter(const unsigned long long* idx, size_t rows, unsigned logK, unsigned bits, fr_t* table,
                                   int sm_count, cudaStream_t s) {
    k_one_hot_scatter<<<grid_for(rows, 256, (size_t)sm_count * 8), 256, 0, s>>>(idx, rows, logK, bits, table);
    return cudaGetLastError();
}
cudaError_t launch_fr_from_u64(const unsigned long long* src, size_t n, fr_t* dst, unsigned bits, int bitrev, int sm_count, cudaStream_t s) {
    k_fr_from_u64<<<grid_for(n, 256, (size_t)sm_count * 8), 256, 0, s>>>(src, n, dst, bits, bitrev);
    return cudaGetLastError();
}

cudaError_t launch_sparse_scatter(const unsigned long long* idx, const fr_t* vals, size_t count, unsigned bits, fr_t* table, int sm_count, cudaStream_t s) {
    k_sparse_scatter<<<grid_for(count, 256, (size_t)sm_count * 8), 256, 0, s>>>(idx, vals, count, bits, table);
    return cudaGetLastError();
}
cudaError_t launch_lt_table(unsigned n, fr_t* out, int sm_count, cudaStream_t s) {
    size_t size = (size_t)1 << (2 * n);
    k_lt_table<<<grid_for(size, 256, (size_t)sm_count * 8), 256, 0, s>>>(n, out, size);
    return cudaGetLastError();
}
cudaError_t launch_table_add(const fr_t* a, const fr_t* b, fr_t* out, size_t n, int sm_count, cudaStream_t s) {
    k_table_add<<<grid_for(n, MLE_THREADS, (size_t)sm_count * 8), MLE_THREADS, 0, s>>>(a, b, out, n);
    return cudaGetLastError();
}
cudaError_t launch_table_scale(const fr_t* a, fr_t* out, size_t n, const fr_t& scalar, int sm_count, cudaStream_t s) {
    k_table_scale<<<grid_for(n, MLE_THREADS, (size_t)sm_count * 8), MLE_THREADS, 0, s>>>(a, out, n, fr_ctab::make(scalar));
    return cudaGetLastError();
}
cudaError_t launch_table_sum(const fr_t* a, size_t n, fr_t* partials, unsigned int* ticket, fr_t* out, int sm_count, cudaStream_t s) {
    k_table_sum<<<grid_for(n, MLE_THREADS, (size_t)sm_count * 4), MLE_THREADS, 0, s>>>(a, n, partials, ticket, out);
    return cudaGetLastError();
}

size_t colsum_splits(size_t rows, size_t cols, int sm_count) {
    // enough blocks to fill the machine ~4x; never more splits than rows/4
    size_t col_blocks = (cols + MLE_THREADS - 1) / MLE_THREADS;
    size_t want = ((size_t)sm_count * 8 + col_blocks - 1) / col_blocks;
    size_t maxs = rows / 4 ? rows / 4 : 1;
    if (want > maxs) want = maxs;
    if (want > 65535) want = 65535;
    return want ? want : 1;
}
cudaError_t launch_colsum(const fr_t* T, const fr_t* W, size_t rows, size_t cols, size_t nsplit, fr_t* partial, fr_t* out, cudaStream_t s) {
    dim3 grid((unsigned)((cols + MLE_THREADS - 1) / MLE_THREADS), (unsigned)nsplit);
    k_colsum<<<grid, MLE_THREADS, 0, s>>>(T, W, rows, cols, nsplit == 1 ? out : partial);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || nsplit == 1) return e;
    k_colsum_finish<<<(unsigned)((cols + 255) / 256), 256, 0, s>>>(partial, nsplit, cols, out);
    return cudaGetLastError();
}
cudaError_t launch_dot(const fr_t* A, const fr_t* B, size_t n, fr_t* partials, unsigned int* ticket, fr_t* out, int sm_count, cudaStream_t s) {
    k_dot<<<grid_for(n, MLE_THREADS, (size_t)sm_count * 2), MLE_THREADS, 0, s>>>(A, B, n, partials, ticket, out);
    return cudaGetLastError();
}

}  // namespace tsg
