// mle.cuh - launch interface of the MLE kernels (mle.cu)
#pragma once
#include <cuda_runtime.h>
#include "fp.cuh"

namespace tsg {

constexpr int MLE_THREADS = 256;

cudaError_t launch_eq_table_bitrev(const fr_t* w_dev, unsigned b, fr_t* out, int sm_count, cudaStream_t s);
cudaError_t launch_bitrev_permute(const fr_t* src, fr_t* dst, unsigned bits, int sm_count, cudaStream_t s);
cudaError_t launch_one_hot_scatter(const unsigned long long* idx, size_t rows, unsigned logK, unsigned bits, fr_t* table,
                                   int sm_count, cudaStream_t s);
cudaError_t launch_fr_from_u64(const unsigned long long* src, size_t n, fr_t* dst, unsigned bits, int bitrev, int sm_count, cudaStream_t s);
cudaError_t launch_sparse_scatter(const unsigned long long* idx, const fr_t* vals, size_t count, unsigned bits, fr_t* table, int sm_count, cudaStream_t s);
cudaError_t launch_lt_table(unsigned n, fr_t* out, int sm_count, cudaStream_t s);   // 2n-variable less-than table, bit-reversed order
cudaError_t launch_table_add(const fr_t* a, const fr_t* b, fr_t* out, size_t n, int sm_count, cudaStream_t s);
cudaError_t launch_table_scale(const fr_t* a, fr_t* out, size_t n, const fr_t& scalar, int sm_count, cudaStream_t s);
cudaError_t launch_table_sum(const fr_t* a, size_t n, fr_t* partials, unsigned int* ticket, fr_t* out, int sm_count, cudaStream_t s);
size_t colsum_splits(size_t rows, size_t cols, int sm_count);
// out[c] = sum_row W[row] * T[row * cols + c]; `partial` must hold nsplit * cols elements when nsplit > 1
cudaError_t launch_colsum(const fr_t* T, const fr_t* W, size_t rows, size_t cols, size_t nsplit, fr_t* partial, fr_t* out, cudaStream_t s);
cudaError_t launch_dot(const fr_t* A, const fr_t* B, size_t n, fr_t* partials, unsigned int* ticket, fr_t* out, int sm_count, cudaStream_t s);

}  // namespace tsg
