// interp.cuh - interface of the device interpolation on {0..n-1} (interp.cu)
#pragma once
#include <cuda_runtime.h>
#include "context.cuh"

namespace tsg {
cudaError_t interp_prepare(tsgpu_ctx* ctx, unsigned logn);
// vals: 2^logn values on the device (natural order); coeffs: 2^logn coefficients low -> high (may alias vals)
cudaError_t interp_run(tsgpu_ctx* ctx, const fr_t* vals, unsigned logn, fr_t* coeffs);
void interp_destroy(tsgpu_ctx* ctx);
}  // namespace tsg
