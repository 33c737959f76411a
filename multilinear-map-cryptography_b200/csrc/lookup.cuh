// lookup.cuh - launch interface of the lookup-argument kernels (lookup.cu)
#pragma once
#include <cuda_runtime.h>
#include "fp.cuh"

namespace tsg {

// acc[8 x + i] += limb_i(W[bitrev_l(j)]) for every j < n, x = idx[j]   (acc: 8 zero-initialised u64 per bucket)
cudaError_t launch_weighted_hist(const fr_t* W, unsigned l, const unsigned long long* idx, size_t n, unsigned long long* acc, int sm_count, cudaStream_t s);
// out[bitrev_k(x)] = sum_i acc[8 x + i] 2^(32 i) mod r, x < 2^k
cudaError_t launch_limb_sums_to_table(const unsigned long long* acc, unsigned k, fr_t* out, int sm_count, cudaStream_t s);
// out[bitrev_l(j)] = j < n ? src[bitrev_k(idx[j])] : 0, j < 2^l
cudaError_t launch_table_gather(const fr_t* src, unsigned k, const unsigned long long* idx, size_t n, unsigned l, fr_t* out, int sm_count, cudaStream_t s);

}  // namespace tsg
