// poly.cuh - launch interface of the univariate open kernels (poly.cu)
#pragma once
#include <cuda_runtime.h>
#include "fp.cuh"

namespace tsg {

constexpr int POLY_THREADS = 256;
constexpr size_t POLY_SPAN = 16;   // coefficients per thread
constexpr int POLY_PW = 8;         // log2(POLY_THREADS) scan weights

inline size_t poly_num_blocks(size_t n) { size_t per = (size_t)POLY_THREADS * POLY_SPAN; return (n + per - 1) / per; }

// pw_dev[s] = z^(POLY_SPAN * 2^s) for s < POLY_PW; W = z^(POLY_THREADS * POLY_SPAN).
// q receives n - 1 coefficients (n >= 1), value one element.
cudaError_t poly_open_launch(const fr_t* c, size_t n, const fr_t& z, const fr_t* pw_dev, const fr_t& W, fr_t* totals, fr_t* carry,
                             fr_t* q, fr_t* value, cudaStream_t s, unsigned* launches);

}  // namespace tsg
