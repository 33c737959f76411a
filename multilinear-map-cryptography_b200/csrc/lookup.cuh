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

// ---- read/write memory (Twist) tables over (cell x, cycle j), reference index x + 2^k j
// Val(x, j): content of cell x just before operation j (2^(k + t) entries, pre-zeroed): write w (address wa[w], operation wj[w], value wv[w]) fills the
// cycles wj[w] + 1 .. wnext[w] - 1 of its cell's row
cudaError_t launch_val_fill(const unsigned long long* wa, const unsigned long long* wj, const unsigned long long* wnext, const fr_t* wv, size_t nwrites,
                            unsigned k, unsigned t, fr_t* out, int sm_count, cudaStream_t s);
// out[addr[j] + 2^k j] = W[j] for j < n with sel[j] == flag (pre-zeroed 2^(k + t) table; W: 2^t entries in table order)
cudaError_t launch_one_hot_weighted(const fr_t* W, const unsigned long long* addr, const unsigned char* sel, unsigned char flag, size_t n, unsigned k, unsigned t,
                                    fr_t* out, int sm_count, cudaStream_t s);
// inout[(x, j)] = rows[j] - inout[(x, j)] over a 2^(k + t)-entry (cell, cycle) table (n entries), rows: 2^t entries in table order
cudaError_t launch_broadcast_rows_minus(const fr_t* rows, unsigned t, fr_t* inout, size_t n, int sm_count, cudaStream_t s);
cudaError_t launch_table_mul(const fr_t* a, const fr_t* b, fr_t* out, size_t n, int sm_count, cudaStream_t s);
// V[a] = LT~(a, b): the less-than indicator [a < c] (natural integer order) extended multilinearly in c and evaluated at the field point b
cudaError_t launch_lt_point_table(const fr_t* b_dev, unsigned t, fr_t* out, int sm_count, cudaStream_t s);

}  // namespace tsg
