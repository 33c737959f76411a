// ntt.cuh - launch interface of the batched NTT kernels (ntt.cu)
#pragma once
#include <cuda_runtime.h>
#include "fp.cuh"

namespace tsg {

constexpr unsigned NTT_LOG_CHUNK = 10;
constexpr size_t NTT_CHUNK = (size_t)1 << NTT_LOG_CHUNK;   // elements per shared-memory chunk (32 KB)
constexpr int NTT_THREADS = 256;

// tw[k] = omega_m^k, twi[k] = omega_m^-k for k < m/2, m = 2^logm
cudaError_t ntt_forward(fr_t* data, unsigned logm, size_t batch, const fr_t* tw, int sm_count, cudaStream_t s, unsigned* launches);
cudaError_t ntt_inverse_unscaled(fr_t* data, unsigned logm, size_t batch, const fr_t* twi, int sm_count, cudaStream_t s, unsigned* launches);

}  // namespace tsg
