// msm.cuh - launch interface of the G1 MSM / SRS kernels (msm.cu)
#pragma once
#include <cuda_runtime.h>
#include "g1.cuh"

namespace tsg {

constexpr unsigned MSM_CHUNK = 64;        // max entries one work item adds into its accumulator (bounds the serial chain of one thread)
constexpr int MSM_ACC_THREADS = 128;
constexpr unsigned MSM_RED_SPAN = 32;     // buckets per thread in the window reduction
constexpr int MSM_SUM_THREADS = 128;
constexpr size_t MSM_POW_SPAN = 64;       // consecutive tau powers per thread
constexpr size_t MSM_INV_SPAN = 32;       // points per batch inversion

struct MsmLayout {
    unsigned c, W, blocks_per_window;
    size_t nbuckets, max_items;
    size_t dig, sorted, hist, offsets, cursor, items, item_off, item_bucket, n_items, order, len_hist, scan_tmp, partial, blockres, window_out;
};

unsigned msm_window_bits(size_t n);
size_t msm_scratch_bytes(size_t n, unsigned c, MsmLayout* L);
// runs steps 1-4; the W window sums (Jacobian) are left at scratch + L.window_out
cudaError_t msm_run(const g1_affine* bases, const fr_t* scalars, size_t n, const MsmLayout& L, unsigned char* scratch, int sm_count,
                    cudaStream_t s, unsigned* launches, cudaEvent_t* acc_events = nullptr);   // acc_events[2]: around the accumulate kernel

cudaError_t launch_tau_powers(const fr_t& tau, size_t first, size_t n, fr_t* out, int sm_count, cudaStream_t s);
cudaError_t launch_fixed_base_mul(const fr_t* scalars, size_t n, const g1_affine* table, g1_xyzz* out, int sm_count, cudaStream_t s);
cudaError_t launch_batch_to_affine(const g1_xyzz* in, size_t n, g1_affine* out, int sm_count, cudaStream_t s);
cudaError_t launch_jac_to_xyzz(const g1_jac* in, size_t n, g1_xyzz* out, int sm_count, cudaStream_t s);
cudaError_t launch_affine_to_jac(const g1_affine* in, size_t n, g1_jac* out, int sm_count, cudaStream_t s);

}  // namespace tsg
