// interp.cuh - interface of the device interpolation on {0..n-1} (interp.cu)
#pragma once
#include <cuda_runtime.h>
#include "context.cuh"

namespace tsg {
cudaError_t interp_prepare(tsgpu_ctx* ctx, unsigned logn);
// vals: 2^logn values on the device (natural order); coeffs: 2^logn coefficients low -> high (may alias vals)
// n_valid (0 = all): interpolate only the first n_valid points (the rest of `vals` must be zero): degree < n_valid, coefficients beyond are zero
cudaError_t interp_run(tsgpu_ctx* ctx, const fr_t* vals, unsigned logn, fr_t* coeffs, size_t n_valid = 0);
void interp_destroy(tsgpu_ctx* ctx);
// *ifact = device table of 1/k! for k = 0 .. 2^logn (at least); owned by the context's interpolation plan
cudaError_t interp_factorials(tsgpu_ctx* ctx, unsigned logn, const fr_t** ifact);
}  // namespace tsg
