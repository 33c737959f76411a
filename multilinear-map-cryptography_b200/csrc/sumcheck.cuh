// sumcheck.cuh - launch interface of the sum-check round kernels (sumcheck.cu)
#pragma once
#include <cuda_runtime.h>
#include "fp.cuh"

namespace tsg {

constexpr int SC_THREADS = 256;
constexpr int SC_BLOCKS_PER_SM = 2;        // evaluation kernels: ~100+ registers per thread
constexpr int SC_BLOCKS_PER_SM_BIND = 4;   // bind-only kernel: light
constexpr int SC_MAX_TABLES = 3;
struct ScTables { fr_t* t[SC_MAX_TABLES]; };

// peer mailboxes of the sharded sum-check (sumcheck.cu "round sums over the ranks"): per rank one device allocation of SC_PEER_MBOX_BYTES,
// region 1: 2 parities x SC_MAX_PEERS slots of SC_PEER_SLOT bytes (round sums: up to 4 field elements + flag),
// region 2: 2 parities x SC_MAX_PEERS slots of SC_PEER_AG_SLOT bytes (small all-gathers: up to SC_PEER_AG_DATA bytes + flag)
constexpr int SC_MAX_PEERS = 8;
constexpr size_t SC_PEER_SLOT = 256, SC_PEER_SLOT_DATA = 128;
constexpr size_t SC_PEER_AG_SLOT = 512, SC_PEER_AG_DATA = 448;
constexpr size_t SC_PEER_MBOX_BYTES = 2 * SC_MAX_PEERS * (SC_PEER_SLOT + SC_PEER_AG_SLOT);
cudaError_t sc_peer_configure(unsigned char* const* mbox, int nranks, int rank, int* host_err, cudaStream_t s);
cudaError_t sc_peer_enable(bool on, cudaStream_t s);     // stream-ordered switch: the round kernels launched after it exchange their sums with the peers
cudaError_t launch_peer_allgather(const void* in, size_t bytes, void* out, cudaStream_t s);

// number of blocks the evaluation kernels may launch (sizes the partial-sum scratch)
inline int sc_max_grid(int sm_count) { return sm_count * SC_BLOCKS_PER_SM_BIND; }

void set_prefetch_min_work(size_t positions);   // d = 2 round kernels with a warp-private shared-memory prefetch of the next tile
// claim (optional, d = 2 only): the value g(0) + g(1) must have; the kernel then sums g(0) and g(2) and derives g(1)
cudaError_t launch_round_eval(int d, const ScTables& tabs, size_t n, fr_t* partials, unsigned int* ticket, fr_t* out4,
                              int sm_count, cudaStream_t s, const fr_t* claim = nullptr);
cudaError_t launch_bind(fr_t* t, size_t n, const fr_t& r, int sm_count, cudaStream_t s);
cudaError_t launch_bind_to(const fr_t* t, fr_t* out, size_t n, const fr_t& r, int sm_count, cudaStream_t s);
// claim (optional): the value g(0) + g(1) of the round being evaluated must have; with it the d = 2 kernel derives g(1) instead of summing it
cudaError_t launch_bind_eval(int d, const ScTables& tabs, size_t n, const fr_t& r, const fr_t* claim, fr_t* partials, unsigned int* ticket,
                             fr_t* out4, int sm_count, cudaStream_t s);

}  // namespace tsg
