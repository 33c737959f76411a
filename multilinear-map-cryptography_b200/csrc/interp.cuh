// interp.cuh - interface of the device interpolation on {0..n-1} (interp.cu)
#pragma once
#include <cuda_runtime.h>
#include "context.cuh"

namespace tsg {
cudaError_t interp_prepare(tsgpu_ctx* ctx, unsigned logn);
// vals: 2^logn values on the device (natural order); coeffs: 2^logn coefficients low -> high (may alias vals)
cudaError_t interp_run(tsgpu_ctx* ctx, const fr_t* vals, unsigned logn, fr_t* coeffs);
void interp_destroy(tsgpu_ctx* ctx);
// *ifact = device table of 1/k! for k = 0 .. 2^logn (at least); owned by the context's interpolation plan
cudaError_t interp_factorials(tsgpu_ctx* ctx, unsigned logn, const fr_t** ifact);
}  // namespace tsg
