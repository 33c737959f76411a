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

// ---- persistent tail of the d = 2 sum-check (sumcheck.cu "persistent tail"): once the two tables fit one CTA's shared memory (<= SC_TAIL_MAX
// entries each after the next fold) ONE resident kernel runs all remaining rounds; it hands the round values to the host and takes the next
// challenge through a mailbox in mapped pinned memory, so a round costs a PCIe round trip instead of a launch + stream synchronisation.
constexpr unsigned SC_TAIL_MAX_LOG = 11;                 // 2 tables x 2^11 entries x 32 B = 128 KiB of shared memory
constexpr unsigned SC_TAIL_ABORT = 0xffffffffu;           // leave at once (the tables are abandoned)
constexpr unsigned SC_TAIL_FLUSH = 0xfffffffeu;           // write the current tables back to HBM and leave (the caller goes on with the per-round kernels)
struct ScTailBox {                                       // pinned + mapped; 64-byte separated fields
    volatile unsigned out_seq; unsigned pad0[15];        // GPU -> host: number of results published so far
    unsigned long long out[16];                          // GPU -> host: g(0), g(2) of the round (raw sums), or the two fully bound table values
    volatile unsigned in_seq; unsigned pad1[15];         // host -> GPU: number of challenges provided so far (SC_TAIL_ABORT: leave)
    uint32_t ctab[64];                                   // host -> GPU: constant-multiplier table of the challenge (fr_ctab)
    volatile unsigned err; unsigned pad2[15];            // GPU -> host: 1 = timed out waiting for the host
};
// launches the tail on tables A, B of n entries (n / 2 <= 2^SC_TAIL_MAX_LOG, n >= 4): folds with r, evaluates the next round and publishes it (out_seq = 1)
cudaError_t launch_sc_tail(fr_t* A, fr_t* B, size_t n, const fr_t& r, ScTailBox* box, cudaStream_t s);
void sc_tail_post_challenge(ScTailBox* box, const fr_t& r, unsigned seq);            // host: next challenge (seq = 1, 2, ...)
void sc_tail_post_command(ScTailBox* box, unsigned command);                         // SC_TAIL_ABORT / SC_TAIL_FLUSH
// host: spin until out_seq == seq (false: time-out, kernel error or GPU-side time-out); values -> out[2]
bool sc_tail_wait(ScTailBox* box, unsigned seq, cudaStream_t s, fr_t out[2]);

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
