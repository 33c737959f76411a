// context.cuh - internal definitions behind the opaque handles of include/tsgpu.h
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <chrono>
#include <cstdint>
#include <map>
#include <string>
#include <vector>
#include "../../include/tsgpu.h"
#include "fp.cuh"
#include "sumcheck.cuh"

struct tsgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;   // side stream of tsgpu_poly_upload_padded_async (created on first use)
    bool h2d_overlap = true;              // Twist::prove from host buffers: the value vector travels while the address vector is committed (tuning "h2d_overlap")
    int sm_count = 0;
    std::string err;
    uint64_t launches = 0;
    // small persistent scratch: block partial sums, ticket counter, 8 result slots, pinned mirror
    tsg::fr_t* partials = nullptr;
    unsigned int* ticket = nullptr;
    tsg::fr_t* dev_out = nullptr;      // 8 elements
    tsg::fr_t* host_out = nullptr;     // pinned, 8 elements
    tsg::fr_t* host_scratch = nullptr; // pinned, 64 elements (host leg of the batch inversion, lagrange.cu)
    unsigned char* host_msm = nullptr; // pinned, HOST_MSM_BYTES: window sums + counters of an MSM pass land here (a pageable target costs a staged copy per transfer)
    static constexpr size_t HOST_MSM_BYTES = 64 << 10;
    tsg::ScTailBox* tail_box = nullptr; // pinned + mapped mailbox of the persistent sum-check tail (sumcheck.cu)
    bool sc_tail = true;               // d = 2 claim-form rounds on tables that fit one CTA's shared memory run in ONE persistent kernel (tuning "sc_tail")
    void* comm = nullptr;              // multi-GPU communicator (comm.cu), optional
    void* interp = nullptr;            // cached interpolation plan (interp.cu)
    // optional per-kernel device timing (CUDA events on `stream`), enabled by tsgpu_set_tuning("kernel_timing", 1)
    // grow-only scratch arenas for the large per-call work buffers (MSM sort/bucket scratch, interpolation ping-pong,
    // quotient) - reused across calls so that the hot path performs no allocation at all
    enum { ARENA_MSM = 0, ARENA_INTERP = 1, ARENA_QUOT = 2, ARENA_BARY = 3, ARENA_COUNT = 4 };
    void* arena[ARENA_COUNT] = {nullptr, nullptr, nullptr, nullptr};
    size_t arena_bytes[ARENA_COUNT] = {0, 0, 0, 0};
    // node inverses 1/(z - j) of the last barycentric opening (lagrange.cu), reused while the opening point stays the same
    void* bary_inv = nullptr; void* bary_spans = nullptr; size_t bary_n = 0; tsgpu_fr bary_z;
    uint64_t msm_points = 0;           // points processed by MSMs (for points/s reporting)
    uint64_t msm_entries = 0;          // bucket entries (non-zero signed digits) = mixed additions of k_msm_accumulate
    uint64_t msm_calls = 0;
    bool timing = false;
    bool msm_tables = true;            // SRS handles carry precomputed window tables (tuning "msm_tables", read when an SRS / basis is built)
    bool eval_basis = true;            // Twist/Shout::prove commit through the Lagrange-basis SRS when it exists (tuning "eval_basis")
    bool deferred_claim_check = false; // opt-in (tuning "deferred_claim_check"): SumCheck::prove, d = 2, runs round 0 in the claim form and checks the claimed sum at the end; default = the reference's deterministic round-0 check before anything is appended (sumcheck.rs:77-84)
    bool peer_exchange = true;         // sharded paths use the peer mailboxes (round sums inside the round kernel, single-kernel all-gathers) when the communicator could map them (tuning "peer_exchange"; 0 = NCCL collectives)
    struct Pending { std::string name; cudaEvent_t a, b; };
    std::vector<Pending> pending;
    std::map<std::string, std::pair<double, uint64_t>> timers;   // name -> (total ms, launches)
};

struct tsgpu_table {
    tsg::fr_t* d = nullptr;   // 2^num_vars entries, bit-reversed index order
    unsigned num_vars = 0;
    size_t capacity = 0;      // allocated entries
};

struct tsgpu_sc {
    tsgpu_ctx* ctx = nullptr;
    int d = 0;
    unsigned vars_left = 0;
    tsgpu_table* tables[tsg::SC_MAX_TABLES] = {nullptr, nullptr, nullptr};
    bool exclusive = false;            // the caller enqueues nothing else on the context until the rounds end (tsgpu_sc_exclusive): the tail kernel may stay resident
    bool tail_active = false;          // the persistent tail kernel is resident and owns the tables
    unsigned tail_seq = 0;             // results consumed from it so far
};

namespace tsg {

int fail(tsgpu_ctx* ctx, int code, const std::string& msg);
int cuda_fail(tsgpu_ctx* ctx, cudaError_t e, const char* what);

#define TSG_CUDA(ctx, call)                                                   \
    do {                                                                      \
        cudaError_t _e = (call);                                              \
        if (_e != cudaSuccess) return ::tsg::cuda_fail((ctx), _e, #call);     \
    } while (0)

// stream-ordered temporary device buffer
struct TempBuf {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    cudaError_t alloc(size_t bytes, cudaStream_t stream) {
        s = stream;
        return cudaMallocAsync(&p, bytes ? bytes : 32, stream);
    }
    ~TempBuf() { if (p) cudaFreeAsync(p, s); }
    template <class T> T* as() { return (T*)p; }
};

int table_alloc(tsgpu_ctx* ctx, unsigned num_vars, tsgpu_table** out);
// persistent scratch of at least `bytes` (contents undefined); nullptr + *err on failure
void* arena_get(tsgpu_ctx* ctx, int slot, size_t bytes, cudaError_t* err);

// RAII scope timing one kernel (or a short sequence) on the context stream when timing is enabled
// It also opens an NVTX range of the same name (header-only nvtx3: a no-op unless a profiler is attached), so `ncu --nvtx
// --nvtx-include "msm_total/"` or an nsys timeline groups the kernels by phase.
struct KernelTimer {
    tsgpu_ctx* ctx; cudaEvent_t a = nullptr, b = nullptr; const char* name;
    struct Nvtx { explicit Nvtx(const char* n) { nvtxRangePushA(n); } ~Nvtx() { nvtxRangePop(); } } nvtx;
    KernelTimer(tsgpu_ctx* c, const char* n) : ctx(c), name(n), nvtx(n) {
        if (!ctx->timing) return;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, ctx->stream);
    }
    ~KernelTimer() {
        if (!a) return;
        cudaEventRecord(b, ctx->stream);
        ctx->pending.push_back({name, a, b});
    }
};
// host wall clock of a phase of a proving path (includes the device work the phase waits for), folded into ctx->timers under `name` when timing is enabled
struct WallTimer {
    tsgpu_ctx* ctx; const char* name; std::chrono::steady_clock::time_point t0;
    WallTimer(tsgpu_ctx* c, const char* n) : ctx(c), name(n), t0(std::chrono::steady_clock::now()) {}
    ~WallTimer() {
        if (!ctx || !ctx->timing) return;
        auto& t = ctx->timers[name];
        t.first += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); t.second += 1;
    }
};
void timers_collect(tsgpu_ctx* ctx);   // after a stream synchronisation: fold finished event pairs into ctx->timers

}  // namespace tsg
