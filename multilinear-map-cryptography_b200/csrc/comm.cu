// comm.cu - one-process-per-GPU sharding of the sum-check inside the library: an NCCL communicator per
// context and the sharded form of SumCheck::prove (src/sumcheck.rs:56-110) for products of MLE tables.
//
// Rank g owns the slice of every table whose reference index has high bits g.  The rounds over the local variables run
// the same kernels as the single-GPU path; after each round kernel the four partial evaluations are widened to 32
// zero-extended 64-bit limbs on the device, summed over the ranks with ONE ncclAllReduce (ncclUint64 / ncclSum: an integer
// sum is exact, a modular one is not an NCCL reduction) on the context stream, copied to the host (256 bytes), carried and
// reduced mod r.  Every rank feeds the same transcript and draws the same challenge.  With one entry per table left per
// rank the G x d values are all-gathered and the last log2 G rounds are finished on every host.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): single-GPU users do not need it, and in a torch process the already
// loaded NCCL is the one used.  The host exchanges the 128-byte unique id however it likes (torch.distributed, MPI, a file).
#include <dlfcn.h>
#include <nccl.h>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "context.cuh"
#include "fr_device.cuh"
#include "sumcheck.cuh"
#include "../host/field64.hpp"
#include "../host/sumcheck_host.hpp"
#include "../host/transcript.hpp"

using namespace tsg;
using namespace tsg::host;

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
    bool load() {
        if (handle) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { err = std::string("cannot load NCCL: ") + dlerror(); return false; }
        auto sym = [&](const char* n) { void* p = dlsym(handle, n); if (!p) err = std::string("NCCL symbol missing: ") + n; return p; };
        GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
        AllGather = (decltype(AllGather))sym("ncclAllGather");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !AllGather || !GetErrorString) { dlclose(handle); handle = nullptr; return false; }
        return true;
    }
};
NcclApi& nccl() { static NcclApi api; return api; }

struct Comm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    unsigned long long* dev = nullptr;     // 32 limb sums, then gather staging: nranks * 3 field elements
    unsigned long long* host = nullptr;    // pinned mirror
    size_t words = 0;
    // peer mailboxes (sumcheck.cu): this rank's allocation and every peer's as mapped into this process through CUDA IPC.  With them the round
    // sums travel inside the round kernel and the small all-gathers are one single-block kernel: no NCCL call on the proving path.
    bool p2p = false;
    unsigned char* mbox = nullptr;
    unsigned char* peer[SC_MAX_PEERS] = {nullptr};
    unsigned long long* ag_in = nullptr;   // pinned: this rank's all-gather payload
    unsigned long long* ag_out = nullptr;  // pinned: nranks payloads
    int* host_err = nullptr;               // pinned: peer time-out flag
};

int nccl_fail(tsgpu_ctx* ctx, ncclResult_t r, const char* what) {
    return fail(ctx, TSGPU_E_PROOF_GENERATION, std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error"));
}
#define TSG_NCCL(ctx, call)                                              \
    do {                                                                 \
        ncclResult_t _r = (call);                                        \
        if (_r != ncclSuccess) return nccl_fail((ctx), _r, #call);       \
    } while (0)

// limbs[8 e + i] = (u64) limb i of element e
__global__ void k_fr_to_limb_sums(const fr_t* in, unsigned n, unsigned long long* limbs) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 8 * n) limbs[t] = in[t >> 3].l[t & 7];
}

Comm* comm_of(tsgpu_ctx* ctx) { return (Comm*)ctx->comm; }
bool use_p2p(tsgpu_ctx* ctx) { Comm* c = comm_of(ctx); return c && c->p2p && ctx->peer_exchange; }

// Mailboxes: allocate, exchange the CUDA IPC handles over the (already working) NCCL communicator, map every peer's allocation.  Any failure
// leaves p2p off and the NCCL collectives in use.
void setup_mailboxes(tsgpu_ctx* ctx, Comm* c) {
    if (c->nranks < 2 || c->nranks > SC_MAX_PEERS) return;
    if (const char* env = getenv("TSGPU_NO_PEER_EXCHANGE")) { if (atoi(env)) return; }
    cudaIpcMemHandle_t mine;
    const int G = c->nranks;
    unsigned char* handles_dev = nullptr;
    std::vector<cudaIpcMemHandle_t> all((size_t)G);
    bool ok = cudaMalloc((void**)&c->mbox, SC_PEER_MBOX_BYTES) == cudaSuccess && cudaMemset(c->mbox, 0, SC_PEER_MBOX_BYTES) == cudaSuccess &&
              cudaDeviceSynchronize() == cudaSuccess && cudaIpcGetMemHandle(&mine, c->mbox) == cudaSuccess &&
              cudaMalloc((void**)&handles_dev, (size_t)(G + 1) * sizeof(mine)) == cudaSuccess;
    // every rank takes part in the all-gather even if its own set-up failed (a zero handle then tells the peers to stay on NCCL)
    if (!ok) memset(&mine, 0, sizeof(mine));
    bool xfer = handles_dev && cudaMemcpyAsync(handles_dev, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
                nccl().AllGather(handles_dev, handles_dev + sizeof(mine), sizeof(mine), ncclUint8, c->comm, ctx->stream) == ncclSuccess &&
                cudaMemcpyAsync(all.data(), handles_dev + sizeof(mine), (size_t)G * sizeof(mine), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
                cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    if (handles_dev) cudaFree(handles_dev);
    static const cudaIpcMemHandle_t zero = {};
    ok = ok && xfer;
    for (int g = 0; g < G && ok; ++g) ok = memcmp(&all[g], &zero, sizeof(zero)) != 0;
    for (int g = 0; g < G && ok; ++g) {
        if (g == c->rank) { c->peer[g] = c->mbox; continue; }
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[g], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
        c->peer[g] = (unsigned char*)p;
    }
    ok = ok && cudaMallocHost((void**)&c->ag_in, SC_PEER_AG_DATA) == cudaSuccess && cudaMallocHost((void**)&c->ag_out, SC_PEER_AG_DATA * G) == cudaSuccess &&
         cudaMallocHost((void**)&c->host_err, sizeof(int)) == cudaSuccess;
    if (ok) { *c->host_err = 0; ok = sc_peer_configure(c->peer, G, c->rank, c->host_err, ctx->stream) == cudaSuccess; }
    // all ranks must agree: one more tiny all-reduce (min) of the outcome
    int flag = ok ? 1 : 0;
    int* flag_dev = nullptr;
    if (cudaMalloc((void**)&flag_dev, sizeof(int)) == cudaSuccess) {
        cudaMemcpyAsync(flag_dev, &flag, sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
        if (nccl().AllReduce(flag_dev, flag_dev, 1, ncclInt32, ncclMin, c->comm, ctx->stream) != ncclSuccess) flag = 0;
        else { cudaMemcpyAsync(&flag, flag_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream); cudaStreamSynchronize(ctx->stream); }
        cudaFree(flag_dev);
    } else flag = 0;
    cudaGetLastError();
    c->p2p = flag == 1;
}
void release_mailboxes(Comm* c) {
    for (int g = 0; g < c->nranks && g < SC_MAX_PEERS; ++g)
        if (c->peer[g] && g != c->rank) cudaIpcCloseMemHandle(c->peer[g]);
    if (c->mbox) cudaFree(c->mbox);
    if (c->ag_in) cudaFreeHost(c->ag_in);
    if (c->ag_out) cudaFreeHost(c->ag_out);
    if (c->host_err) cudaFreeHost(c->host_err);
}
int peer_timeout(tsgpu_ctx* ctx, Comm* c) {
    if (*c->host_err) { *c->host_err = 0; return fail(ctx, TSGPU_E_PROOF_GENERATION, "peer exchange timed out: a rank of the communicator did not answer"); }
    return TSGPU_OK;
}

// exact field sum over the ranks of `n` (<= 4) elements at ctx->dev_out; result to `out` on the host of every rank
int allreduce_dev_out(tsgpu_ctx* ctx, unsigned n, fr_t* out) {
    Comm* c = comm_of(ctx);
    if (!c || c->nranks == 1) {
        TSG_CUDA(ctx, cudaMemcpyAsync(ctx->host_out, ctx->dev_out, n * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
        TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        memcpy(out, ctx->host_out, n * sizeof(fr_t));
        return TSGPU_OK;
    }
    k_fr_to_limb_sums<<<1, 32, 0, ctx->stream>>>(ctx->dev_out, n, c->dev);
    ctx->launches += 1;
    TSG_NCCL(ctx, nccl().AllReduce(c->dev, c->dev, 8 * n, ncclUint64, ncclSum, c->comm, ctx->stream));
    TSG_CUDA(ctx, cudaMemcpyAsync(c->host, c->dev, 8 * n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    tsgpu_fr tmp[4];
    tsgpu_fr_from_limb_sums((const uint64_t*)c->host, n, tmp);
    memcpy(out, tmp, n * sizeof(fr_t));
    return TSGPU_OK;
}

}  // namespace

extern "C" {

// rank 0 creates the id; the host program hands the 128 bytes to every rank
int tsgpu_comm_unique_id(uint8_t out[128]) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (!out || !nccl().load()) return TSGPU_E_PROOF_GENERATION;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != ncclSuccess) return TSGPU_E_PROOF_GENERATION;
    memcpy(out, &id, 128);
    return TSGPU_OK;
}

int tsgpu_comm_init(tsgpu_ctx* ctx, int nranks, int rank, const uint8_t id[128]) {
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "bad communicator arguments");
    if (nranks & (nranks - 1)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "number of ranks must be a power of two");
    if (ctx->comm) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "communicator already initialised");
    Comm* c = new (std::nothrow) Comm;
    if (!c) return fail(ctx, TSGPU_E_PROOF_GENERATION, "out of host memory");
    c->nranks = nranks; c->rank = rank;
    c->words = 32 + (size_t)nranks * 3 * 4 + 16;
    if (nranks > 1) {
        if (!nccl().load()) { delete c; return fail(ctx, TSGPU_E_PROOF_GENERATION, nccl().err); }
        ncclUniqueId uid; memcpy(&uid, id, 128);
        cudaSetDevice(ctx->device);
        ncclResult_t r = nccl().CommInitRank(&c->comm, nranks, uid, rank);
        if (r != ncclSuccess) { delete c; return nccl_fail(ctx, r, "ncclCommInitRank"); }
    }
    if (cudaMalloc((void**)&c->dev, c->words * 8) != cudaSuccess || cudaMallocHost((void**)&c->host, c->words * 8) != cudaSuccess) {
        cudaGetLastError();
        if (c->comm) nccl().CommDestroy(c->comm);
        if (c->dev) cudaFree(c->dev);
        delete c;
        return fail(ctx, TSGPU_E_PROOF_GENERATION, "communicator buffers");
    }
    ctx->comm = c;
    if (nranks > 1) setup_mailboxes(ctx, c);
    return TSGPU_OK;
}
int tsgpu_comm_peer_exchange(const tsgpu_ctx* ctx) { return ctx && ctx->comm && ((Comm*)ctx->comm)->p2p && ctx->peer_exchange ? 1 : 0; }
int tsgpu_comm_size(const tsgpu_ctx* ctx) { return ctx && ctx->comm ? ((Comm*)ctx->comm)->nranks : 1; }
int tsgpu_comm_rank(const tsgpu_ctx* ctx) { return ctx && ctx->comm ? ((Comm*)ctx->comm)->rank : 0; }
void tsgpu_comm_destroy(tsgpu_ctx* ctx) {
    if (!ctx || !ctx->comm) return;
    Comm* c = (Comm*)ctx->comm;
    cudaStreamSynchronize(ctx->stream);
    release_mailboxes(c);
    if (c->comm) nccl().CommDestroy(c->comm);
    if (c->dev) cudaFree(c->dev);
    if (c->host) cudaFreeHost(c->host);
    delete c;
    ctx->comm = nullptr;
}

// all-gather of `bytes` (a multiple of 8, at most 96 * 4) per rank from and to host memory: partial MSM results (one G1Projective per
// rank, summed by the caller with tsgpu_g1_add - group addition is not an NCCL reduction)
int tsgpu_comm_allgather(tsgpu_ctx* ctx, const void* in, size_t bytes, void* out) {
    if (!ctx || !in || !out || bytes % 8) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "bad all-gather arguments");
    Comm* c = comm_of(ctx);
    if (!c || c->nranks == 1) { memcpy(out, in, bytes); return TSGPU_OK; }
    if (use_p2p(ctx) && bytes <= SC_PEER_AG_DATA) {
        // one single-block kernel: payload from pinned memory into every peer's mailbox, flags, gathered payloads back into pinned memory
        memcpy(c->ag_in, in, bytes);
        TSG_CUDA(ctx, launch_peer_allgather(c->ag_in, bytes, c->ag_out, ctx->stream));
        ctx->launches += 1;
        TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        int rc = peer_timeout(ctx, c);
        if (rc) return rc;
        memcpy(out, c->ag_out, bytes * c->nranks);
        return TSGPU_OK;
    }
    const size_t w = bytes / 8;
    TempBuf s, r;
    TSG_CUDA(ctx, s.alloc(bytes, ctx->stream));
    TSG_CUDA(ctx, r.alloc(bytes * c->nranks, ctx->stream));
    TSG_CUDA(ctx, cudaMemcpyAsync(s.p, in, bytes, cudaMemcpyHostToDevice, ctx->stream));
    TSG_NCCL(ctx, nccl().AllGather(s.p, r.p, w, ncclUint64, c->comm, ctx->stream));
    TSG_CUDA(ctx, cudaMemcpyAsync(out, r.p, bytes * c->nranks, cudaMemcpyDeviceToHost, ctx->stream));
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TSGPU_OK;
}

// MultilinearExtension::evaluate (src/polynomials.rs:85-122) of a table sliced over the ranks: rank g holds the entries whose index has
// high bits g (num_vars - log2 G local variables).  evaluate(point) = sum_g eq(point_high, g) * local_g.evaluate(point_low): one local
// streaming pass, one scalar weight on the host, one exact all-reduce of a single field element.  Every rank gets the value.
int tsgpu_table_evaluate_sharded(tsgpu_ctx* ctx, const tsgpu_table* local, unsigned num_vars, const tsgpu_fr* point, tsgpu_fr* out) {
    if (!ctx || !local || !point || !out) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    Comm* c = comm_of(ctx);
    const int G = c ? c->nranks : 1, rank = c ? c->rank : 0;
    unsigned logG = 0; while ((1 << logG) < G) ++logG;
    if (num_vars < logG || local->num_vars != num_vars - logG) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Number of variables must match");
    const unsigned n_local = num_vars - logG;
    tsgpu_fr v;
    int rc = tsgpu_table_evaluate(ctx, local, point, &v);          // uses point[0 .. n_local)
    if (rc) return rc;
    fr_t acc; memcpy(acc.l, v.l, 32);
    for (unsigned k = 0; k < logG; ++k) {
        fr_t p; memcpy(p.l, point[n_local + k].l, 32);
        acc = acc * (((rank >> k) & 1) ? p : fr_t::one() - p);
    }
    memcpy(ctx->host_out, acc.l, 32);
    TSG_CUDA(ctx, cudaMemcpyAsync(ctx->dev_out, ctx->host_out, sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    fr_t total;
    if ((rc = allreduce_dev_out(ctx, 1, &total))) return rc;
    memcpy(out->l, total.l, 32);
    return TSGPU_OK;
}

// SumCheck::new(num_vars, claimed_sum).prove(|v| prod_t mle_t.evaluate(v), transcript) with the hypercube sliced over the ranks of
// the context's communicator.  tables: this rank's slices (num_vars - log2 G variables each, consumed).  Every rank returns the same
// round polynomials (num_vars x 4), final evaluation, challenges (num_vars) and fully bound table values (d).
int tsgpu_sumcheck_prove_product_sharded(tsgpu_ctx* ctx, tsgpu_table* const* tables, int d, unsigned num_vars, const tsgpu_fr* claimed_sum,
                                         tsgpu_transcript* transcript, tsgpu_fr* round_polys, tsgpu_fr* final_evaluation,
                                         tsgpu_fr* challenges, tsgpu_fr* table_finals) {
    if (!ctx || !tables || !claimed_sum || !transcript || !round_polys || !final_evaluation) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (d < 1 || d > SC_MAX_TABLES) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "sum-check supports products of 1..3 tables");
    Comm* c = comm_of(ctx);
    const int G = c ? c->nranks : 1;
    unsigned logG = 0; while ((1 << logG) < G) ++logG;
    if (num_vars < logG) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "more ranks than table entries");
    const unsigned n_local = num_vars - logG;
    for (int t = 0; t < d; ++t)
        if (!tables[t] || tables[t]->num_vars != n_local) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Number of variables must match");
    Transcript& tr = *tsgpu_transcript_inner(transcript);
    ScTables tabs; for (int i = 0; i < SC_MAX_TABLES; ++i) tabs.t[i] = i < d ? tables[i]->d : nullptr;
    fr_t current; memcpy(current.l, claimed_sum->l, 32);
    std::vector<fr_t> chal;
    // d = 2: round 0 also in the claim form, the claimed sum checked later (host/sumcheck_host.cpp explains why this is the reference's round-0
    // check in effect): with a wrong claim the first FULL evaluation - a tail round, or the final product - disagrees with the running sum
    const bool deferred = d == 2 && n_local > 0 && ctx->deferred_claim_check;
    const size_t transcript_mark = tr.state_len();
    auto claim_failed = [&]() -> int {
        tr.truncate(transcript_mark);
        memset(round_polys, 0, (size_t)num_vars * 4 * sizeof(tsgpu_fr));       // nothing of the aborted run is left in the caller's buffers
        return fail(ctx, TSGPU_E_SUMCHECK, "Round 0 consistency check failed");
    };
    auto absorb = [&](unsigned round, const fr_t e[4], fr_t* r_out) -> int {
        fr_t coeffs[4];
        interpolate4(e, coeffs);
        fr_t g0 = horner_eval(coeffs, 4, fr_t::zero()), g1 = horner_eval(coeffs, 4, fr_t::one());
        if (g0 + g1 != current) return deferred ? claim_failed() : fail(ctx, TSGPU_E_SUMCHECK, "Round " + std::to_string(round) + " consistency check failed");   // sumcheck.rs:77-84
        for (int k = 0; k < 4; ++k) memcpy(round_polys[4 * round + k].l, coeffs[k].l, 32);
        tr.append_field_elements("sumcheck_round_" + std::to_string(round), coeffs, 4);
        fr_t r = tr.challenge_field_element("sumcheck_challenge_" + std::to_string(round));
        chal.push_back(r);
        current = horner_eval(coeffs, 4, r);
        *r_out = r;
        return TSGPU_OK;
    };
    int rc;
    // With peer mailboxes the finishing thread of every round kernel exchanges its sums with the other ranks itself and writes the GLOBAL values to
    // the pinned mirror: a round is one launch + one stream synchronisation.  Otherwise: kernel -> widen -> ncclAllReduce -> copy (allreduce_dev_out).
    const bool p2p = use_p2p(ctx) && G > 1;
    struct PeerScope {          // the switch is stream-ordered: on before the first round kernel, off after the last on every exit path
        tsgpu_ctx* ctx; bool on;
        ~PeerScope() { if (on) sc_peer_enable(false, ctx->stream); }
    } scope{ctx, p2p};
    if (p2p) TSG_CUDA(ctx, sc_peer_enable(true, ctx->stream));
    fr_t* round_out = p2p ? ctx->host_out : ctx->dev_out;
    auto round_values = [&](fr_t* ev) -> int {
        if (!p2p) return allreduce_dev_out(ctx, 4, ev);
        TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        int prc = peer_timeout(ctx, c);
        if (prc) return prc;
        memcpy(ev, ctx->host_out, 4 * sizeof(fr_t));
        return TSGPU_OK;
    };
    // ---- rounds over the local variables
    fr_t ev[4];
    if (n_local) {
        const fr_t zero = fr_t::zero();
        TSG_CUDA(ctx, launch_round_eval(d, tabs, (size_t)1 << n_local, ctx->partials, ctx->ticket, round_out, ctx->sm_count, ctx->stream, deferred ? &zero : nullptr));
        ctx->launches += 1;
        if ((rc = round_values(ev))) return rc;
        if (deferred) {                       // g(1) from the GLOBAL claim after the all-reduce, as in the later rounds
            ev[1] = current - ev[0];
            fr_t dd = ev[2] - ev[1];
            ev[3] = ev[0] + dd + dd + dd;
        }
    }
    // persistent tail (sumcheck.cu): d = 2 rounds on slices that fit one CTA's shared memory run inside ONE resident kernel, which also sums the round values
    // with the peers; the host feeds it the challenges through the mapped mailbox.  Needs the in-kernel exchange (or a single rank).
    const bool tail_ok = d == 2 && ctx->sc_tail && ctx->tail_box && (p2p || G == 1);
    bool tail_active = false; unsigned tail_seq = 0;
    struct TailGuard {          // an error return must not leave the resident kernel behind
        tsgpu_ctx* ctx; bool* active;
        ~TailGuard() { if (*active) { sc_tail_post_command(ctx->tail_box, SC_TAIL_ABORT); cudaStreamSynchronize(ctx->stream); } }
    } tail_guard{ctx, &tail_active};
    auto tail_step = [&](const fr_t& r, size_t n, fr_t v[2]) -> int {
        if (!tail_active) {
            TSG_CUDA(ctx, launch_sc_tail(tables[0]->d, tables[1]->d, n, r, ctx->tail_box, ctx->stream));
            ctx->launches += 1;
            tail_active = true; tail_seq = 0;
        } else sc_tail_post_challenge(ctx->tail_box, r, tail_seq);
        tail_seq += 1;
        if (!sc_tail_wait(ctx->tail_box, tail_seq, ctx->stream, v)) return fail(ctx, TSGPU_E_PROOF_GENERATION, "sum-check tail kernel did not answer");
        return p2p ? peer_timeout(ctx, c) : TSGPU_OK;
    };
    for (unsigned round = 0; round < n_local; ++round) {
        fr_t r;
        if ((rc = absorb(round, ev, &r))) return rc;
        const size_t n = (size_t)1 << (n_local - round);
        if (tail_ok && (tail_active || (n >= 4 && n / 2 <= ((size_t)1 << SC_TAIL_MAX_LOG)))) {
            fr_t v[2];
            if ((rc = tail_step(r, n, v))) return rc;
            if (round + 1 < n_local) {          // v = GLOBAL g(0), g(2) of the next round; g(1) from the global claim
                ev[0] = v[0]; ev[2] = v[1];
                ev[1] = current - ev[0];
                fr_t dd = ev[2] - ev[1];
                ev[3] = ev[0] + dd + dd + dd;
            } else {                            // last local fold done: the kernel has written the bound values to the tables and leaves
                tail_active = false;
                TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            }
            for (int t = 0; t < d; ++t) tables[t]->num_vars -= 1;
            continue;
        }
        if (round + 1 < n_local) {
            // d = 2: the kernel sums g(0) and g(2) only (claim form with claim 0 leaves g(1) slot = -g(0), unused); g(1) follows from
            // the GLOBAL claim after the all-reduce: g(1) = current - g(0), g(3) = g(0) - 3 g(1) + 3 g(2)
            const fr_t zero = fr_t::zero();
            TSG_CUDA(ctx, launch_bind_eval(d, tabs, n, r, d == 2 ? &zero : nullptr, ctx->partials, ctx->ticket, round_out, ctx->sm_count, ctx->stream));
            ctx->launches += 1;
            if ((rc = round_values(ev))) return rc;
            if (d == 2) {
                ev[1] = current - ev[0];
                fr_t dd = ev[2] - ev[1];
                ev[3] = ev[0] + dd + dd + dd;
            }
        } else {
            for (int t = 0; t < d; ++t) { TSG_CUDA(ctx, launch_bind(tables[t]->d, n, r, ctx->sm_count, ctx->stream)); ctx->launches += 1; }
        }
        for (int t = 0; t < d; ++t) tables[t]->num_vars -= 1;
    }
    if (p2p) { scope.on = false; TSG_CUDA(ctx, sc_peer_enable(false, ctx->stream)); }
    // ---- one entry per table per rank: gather (rank = high index bits) and finish the last log2 G rounds on every host
    std::vector<fr_t> tail((size_t)G * d);
    {
        std::vector<fr_t> mine((size_t)d);
        for (int t = 0; t < d; ++t) TSG_CUDA(ctx, cudaMemcpyAsync(ctx->dev_out + t, tables[t]->d, sizeof(fr_t), cudaMemcpyDeviceToDevice, ctx->stream));
        TSG_CUDA(ctx, cudaMemcpyAsync(ctx->host_out, ctx->dev_out, d * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
        TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        memcpy(mine.data(), ctx->host_out, d * sizeof(fr_t));
        if ((rc = tsgpu_comm_allgather(ctx, mine.data(), d * sizeof(fr_t), tail.data()))) return rc;   // tail[g * d + t]
    }
    std::vector<std::vector<fr_t>> T((size_t)d, std::vector<fr_t>((size_t)G));
    for (int g = 0; g < G; ++g) for (int t = 0; t < d; ++t) T[t][g] = tail[(size_t)g * d + t];
    size_t len = (size_t)G;
    for (unsigned k = 0; k < logG; ++k) {
        // round over variable n_local + k = bit k of the rank index: pairs (2 j, 2 j + 1) in natural order
        fr_t e[4] = {fr_t::zero(), fr_t::zero(), fr_t::zero(), fr_t::zero()};
        for (size_t j = 0; j < len / 2; ++j) {
            fr_t v[SC_MAX_TABLES], dlt[SC_MAX_TABLES];
            for (int t = 0; t < d; ++t) { v[t] = T[t][2 * j]; dlt[t] = T[t][2 * j + 1] - T[t][2 * j]; }
            for (int x = 0; x < 4; ++x) {
                fr_t p = v[0];
                for (int t = 1; t < d; ++t) p = p * v[t];
                e[x] = e[x] + p;
                for (int t = 0; t < d; ++t) v[t] = v[t] + dlt[t];
            }
        }
        fr_t r;
        if ((rc = absorb(n_local + k, e, &r))) return rc;
        for (size_t j = 0; j < len / 2; ++j)
            for (int t = 0; t < d; ++t) T[t][j] = T[t][2 * j] + r * (T[t][2 * j + 1] - T[t][2 * j]);
        len /= 2;
    }
    fr_t fe = fr_t::one();
    for (int t = 0; t < d; ++t) { fe = fe * T[t][0]; if (table_finals) memcpy(table_finals[t].l, T[t][0].l, 32); }
    if (deferred && fe != current) return claim_failed();                     // the deferred round-0 check when no tail round ran a full evaluation
    memcpy(final_evaluation->l, fe.l, 32);                                    // polynomial(&fixed_variables), sumcheck.rs:104
    if (challenges) for (size_t i = 0; i < chal.size(); ++i) memcpy(challenges[i].l, chal[i].l, 32);
    return TSGPU_OK;
}

}  // extern "C"
