// msm.cuh - launch interface of the G1 MSM / SRS kernels (msm.cu)
#pragma once
#include <cuda_runtime.h>
#include "g1.cuh"

namespace tsg {

#ifndef TSG_MSM_CHUNK
#define TSG_MSM_CHUNK 64
#endif
constexpr unsigned MSM_CHUNK = TSG_MSM_CHUNK;       // LARGEST number of entries one work item adds into its accumulator (bounds the serial chain of one thread; sizes the length histogram);
                                                    // the length a pass uses (16, 32 or 64) is chosen in msm_scratch_bytes
constexpr unsigned MSM_SERIAL_MERGE = 16;  // passes whose buckets split into at most this many chunks merge them with one thread per bucket (k_msm_merge_serial)
constexpr unsigned MSM_BLOCK_MERGE = 2048; // ... up to this many chunks one block per bucket (k_msm_merge_heavy)
constexpr size_t MSM_ITEMS_TARGET = 240000; // ~3 resident waves of accumulation threads (148 SMs x 512): below it a pass shortens its chunks (msm_scratch_bytes)
constexpr int MSM_ACC_THREADS = 128;
constexpr unsigned MSM_RED_SPAN = 32;     // most buckets per thread in the window reduction (2 x span additions deep; the bit-decomposed tail costs
                                          // log2(buckets / span) / 2 additions per span); small bucket sets use shorter spans to keep the chain short
constexpr size_t MSM_RED_MIN_SPANS = 32768; // ... i.e. halve the span while fewer span threads than this would run (measured with the quad tree, round 2: 8 / 65536 -> 32 / 32768
                                          // takes the reduction of the open pass of a 2^20-op proof from 0.80 to 0.70 ms, of the commit pass from 0.25 to 0.19 ms)
constexpr int MSM_SUM_THREADS = 256;
constexpr int MSM_FIN_THREADS = 64;       // k_msm_bit_finish: 16 lane quads add the <= 32 slice sums of one (set, slot)
constexpr size_t MSM_POW_SPAN = 64;       // consecutive tau powers per thread
constexpr size_t MSM_INV_SPAN = 32;       // points per batch inversion

constexpr int MSM_MAX_BATCH = 4;          // independent MSMs processed in one pass

struct MsmJob {
    const g1_affine* bases;   // n points, or (precomputed-table mode) W tables of `stride` points: table[w * stride + i] = 2^(c w) * P_i
    size_t stride;
    const fr_t* scalars;      // n scalars, Montgomery form
    size_t n;
};
struct MsmBases { const g1_affine* p[MSM_MAX_BATCH]; };

struct MsmLayout {
    unsigned c, W, K, sets, blocks_per_window, span, span_bits, chunk;   // chunk: entries per work item (<= MSM_CHUNK)   // sets = bucket sets = K * (shared ? 1 : W); blocks_per_window = spans per set = 2^span_bits
    bool shared;                                                 // all W digit positions of a job feed one bucket set (needs the precomputed tables)
    size_t nmax, nbuckets, max_items;
    size_t dig, sorted, hist, offsets, cursor, items, item_off, item_bucket, n_items, order, len_hist, scan_tmp, partial, blockres, bits, window_out;
    size_t heavy;        // list of the buckets k_msm_merge_heavy handles
    size_t zero_bytes;   // hist, cursor, n_items, len_hist: one contiguous block cleared at the start of a pass
};

void msm_set_quad_tree(int on);
unsigned msm_window_bits(size_t n);
unsigned msm_table_window_bits(size_t n);
size_t msm_scratch_bytes(size_t nmax, int K, unsigned c, bool shared, MsmLayout* L, unsigned windows = 0);   // windows > 0: scan only that many low digit positions
// runs every device phase; per bucket set (job-major) span_bits + 2 Jacobian points are left at scratch + L.window_out:
// S_w = 2^log2(span) * sum_k 2^k P[k] + P[span_bits] + P[span_bits + 1]  (finished on the host: a Horner pass of ~40 group operations)
cudaError_t msm_run(const MsmJob* jobs, int K, const MsmLayout& L, unsigned char* scratch, int sm_count,
                    cudaStream_t s, unsigned* launches, cudaEvent_t* ev = nullptr);   // ev[5]: start | sort done | accumulate done | chunk merge done | reduce done
// table[w * n + i] = 2^(c w) * bases[i] for w < ceil(255 / c); cur: n XYZZ points of scratch
cudaError_t msm_build_table(const g1_affine* bases, size_t n, unsigned c, g1_affine* table, g1_xyzz* cur, int sm_count, cudaStream_t s, unsigned* launches,
                            unsigned windows = 0);
// *flag (zeroed by the caller) becomes non-zero iff some scalar does not fit 64 bits
cudaError_t launch_scalar_probe(const fr_t* scalars, size_t n, unsigned* flag, int sm_count, cudaStream_t s);
constexpr unsigned MSM_SHORT_C = 17;        // table mode for scalars below 2^64: 17-bit windows (16-bit addresses stay positive digits: no carry bucket) ...
constexpr unsigned MSM_SHORT_WINDOWS = 4;   // ... of which four cover 64 bits plus the signed-digit carry

cudaError_t launch_tau_powers(const fr_t& tau, size_t first, size_t n, fr_t* out, int sm_count, cudaStream_t s);
cudaError_t launch_fixed_base_mul(const fr_t* scalars, size_t n, const g1_affine* table, g1_xyzz* out, int sm_count, cudaStream_t s);
cudaError_t launch_batch_to_affine(const g1_xyzz* in, size_t n, g1_affine* out, int sm_count, cudaStream_t s);
cudaError_t launch_jac_to_xyzz(const g1_jac* in, size_t n, g1_xyzz* out, int sm_count, cudaStream_t s);
cudaError_t launch_affine_to_jac(const g1_affine* in, size_t n, g1_jac* out, int sm_count, cudaStream_t s);

}  // namespace tsg
