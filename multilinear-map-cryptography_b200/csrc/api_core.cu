// api_core.cu - C ABI (include/tsgpu.h): context, MLE tables, evaluate / partial evaluate, sum-check rounds.
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <set>
#include <new>
#include <string>
#include <vector>
#include "context.cuh"
#include "mle.cuh"
#include "sumcheck.cuh"
#include "interp.cuh"
#include "lookup.cuh"
#include "msm.cuh"

using namespace tsg;

namespace tsg {

int fail(tsgpu_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}
int cuda_fail(tsgpu_ctx* ctx, cudaError_t e, const char* what) {
    // CUDA failures surface as TwistAndShoutError::ProofGeneration
    std::string m = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    cudaGetLastError();
    return fail(ctx, TSGPU_E_PROOF_GENERATION, m);
}

void timers_collect(tsgpu_ctx* ctx) {
    std::set<cudaEvent_t> events;   // consecutive phases share their boundary event: destroy each once, after all pairs are read
    for (auto& p : ctx->pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { auto& t = ctx->timers[p.name]; t.first += ms; t.second += 1; }
        events.insert(p.a); events.insert(p.b);
    }
    for (cudaEvent_t e : events) cudaEventDestroy(e);
    ctx->pending.clear();
    cudaGetLastError();
}

void* arena_get(tsgpu_ctx* ctx, int slot, size_t bytes, cudaError_t* err) {
    *err = cudaSuccess;
    if (ctx->arena_bytes[slot] >= bytes && ctx->arena[slot]) return ctx->arena[slot];
    cudaStreamSynchronize(ctx->stream);
    if (ctx->arena[slot]) cudaFree(ctx->arena[slot]);
    ctx->arena[slot] = nullptr; ctx->arena_bytes[slot] = 0;
    size_t want = bytes + bytes / 8 + (1 << 20);   // head-room so that slightly larger requests do not reallocate
    *err = cudaMalloc(&ctx->arena[slot], want);
    if (*err != cudaSuccess) return nullptr;
    ctx->arena_bytes[slot] = want;
    return ctx->arena[slot];
}

int table_alloc(tsgpu_ctx* ctx, unsigned num_vars, tsgpu_table** out) {
    if (num_vars > 34) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "num_vars too large");
    tsgpu_table* t = new (std::nothrow) tsgpu_table;
    if (!t) return fail(ctx, TSGPU_E_PROOF_GENERATION, "out of host memory");
    t->num_vars = num_vars;
    t->capacity = (size_t)1 << num_vars;
    cudaError_t e = cudaMallocAsync((void**)&t->d, t->capacity * sizeof(fr_t), ctx->stream);
    if (e != cudaSuccess) { delete t; return cuda_fail(ctx, e, "cudaMallocAsync(table)"); }
    *out = t;
    return TSGPU_OK;
}

}  // namespace tsg

static_assert(sizeof(fr_t) == 32 && sizeof(tsgpu_fr) == 32, "Fr element must be 32 bytes");

static inline fr_t to_fr(const tsgpu_fr* p) { fr_t r; memcpy(r.l, p->l, 32); return r; }

extern "C" {

int tsgpu_abi_version(void) { return TSGPU_ABI_VERSION; }

int tsgpu_init(int device, void* stream, tsgpu_ctx** out) {
    if (!out) return TSGPU_E_INVALID_PARAMETERS;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return TSGPU_E_PROOF_GENERATION;   // no CUDA device: there is no CPU fallback
    }
    tsgpu_ctx* ctx = new (std::nothrow) tsgpu_ctx;
    if (!ctx) return TSGPU_E_PROOF_GENERATION;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return TSGPU_E_PROOF_GENERATION; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return TSGPU_E_PROOF_GENERATION; }
    ctx->sm_count = prop.multiProcessorCount;
    if (stream) { ctx->stream = (cudaStream_t)stream; ctx->own_stream = false; }
    else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return TSGPU_E_PROOF_GENERATION; }
        ctx->own_stream = true;
    }
    // keep freed stream-ordered allocations cached in the pool instead of returning them to the OS
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thresh = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
    }
    size_t np = (size_t)sc_max_grid(ctx->sm_count) * 4 + 64;
    bool ok = cudaMalloc((void**)&ctx->partials, np * sizeof(fr_t)) == cudaSuccess &&
              cudaMalloc((void**)&ctx->ticket, 64) == cudaSuccess &&
              cudaMalloc((void**)&ctx->dev_out, 8 * sizeof(fr_t)) == cudaSuccess &&
              cudaMallocHost((void**)&ctx->host_out, 8 * sizeof(fr_t)) == cudaSuccess &&
              cudaMallocHost((void**)&ctx->host_scratch, 64 * sizeof(fr_t)) == cudaSuccess &&
              cudaMallocHost((void**)&ctx->host_msm, tsgpu_ctx::HOST_MSM_BYTES) == cudaSuccess &&
              cudaHostAlloc((void**)&ctx->tail_box, sizeof(ScTailBox), cudaHostAllocMapped) == cudaSuccess &&
              cudaMemset(ctx->ticket, 0, 64) == cudaSuccess;
    if (!ok) { cudaGetLastError(); tsgpu_destroy(ctx); return TSGPU_E_PROOF_GENERATION; }
    *out = ctx;
    return TSGPU_OK;
}

void tsgpu_destroy(tsgpu_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    tsgpu_comm_destroy(ctx);
    interp_destroy(ctx);
    for (int i = 0; i < tsgpu_ctx::ARENA_COUNT; ++i) if (ctx->arena[i]) cudaFree(ctx->arena[i]);
    if (ctx->partials) cudaFree(ctx->partials);
    if (ctx->ticket) cudaFree(ctx->ticket);
    if (ctx->dev_out) cudaFree(ctx->dev_out);
    if (ctx->host_out) cudaFreeHost(ctx->host_out);
    if (ctx->host_scratch) cudaFreeHost(ctx->host_scratch);
    if (ctx->host_msm) cudaFreeHost(ctx->host_msm);
    if (ctx->tail_box) cudaFreeHost(ctx->tail_box);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* tsgpu_last_error(const tsgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }
uint64_t tsgpu_launch_count(const tsgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }
int tsgpu_sm_count(const tsgpu_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
// monotonically increasing work counters: "launches", "msm_calls", "msm_points", "msm_entries" (bucket entries = mixed additions)
uint64_t tsgpu_counter_read(const tsgpu_ctx* ctx, const char* name) {
    if (!ctx || !name) return 0;
    if (!strcmp(name, "launches")) return ctx->launches;
    if (!strcmp(name, "msm_calls")) return ctx->msm_calls;
    if (!strcmp(name, "msm_points")) return ctx->msm_points;
    if (!strcmp(name, "msm_entries")) return ctx->msm_entries;
    return 0;
}
int tsgpu_synchronize(tsgpu_ctx* ctx) {
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TSGPU_OK;
}

int tsgpu_timer_read(tsgpu_ctx* ctx, const char* name, double* total_ms, uint64_t* count) {
    if (!ctx || !name) return TSGPU_E_INVALID_PARAMETERS;
    cudaStreamSynchronize(ctx->stream);
    timers_collect(ctx);
    auto it = ctx->timers.find(name);
    if (total_ms) *total_ms = it == ctx->timers.end() ? 0.0 : it->second.first;
    if (count) *count = it == ctx->timers.end() ? 0 : it->second.second;
    return TSGPU_OK;
}
void tsgpu_timer_reset(tsgpu_ctx* ctx) {
    if (!ctx) return;
    cudaStreamSynchronize(ctx->stream);
    timers_collect(ctx);
    ctx->timers.clear();
}

int tsgpu_set_tuning(tsgpu_ctx* ctx, const char* key, long value) {
    if (!key) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null key");
    if (!strcmp(key, "prefetch_min_log2")) {   // d = 2 rounds with >= 2^value positions per launch use the warp-private prefetch kernels; < 0 disables
        set_prefetch_min_work(value < 0 || value > 61 ? (size_t)1 << 62 : (size_t)1 << value);
        return TSGPU_OK;
    }
    if (!strcmp(key, "sc_tail")) { ctx->sc_tail = value != 0; return TSGPU_OK; }
    if (!strcmp(key, "peer_exchange")) { ctx->peer_exchange = value != 0; return TSGPU_OK; }
    if (!strcmp(key, "deferred_claim_check")) { ctx->deferred_claim_check = value != 0; return TSGPU_OK; }
    if (!strcmp(key, "kernel_timing")) { ctx->timing = value != 0; return TSGPU_OK; }
    if (!strcmp(key, "msm_quad_tree")) { tsg::msm_set_quad_tree(value != 0); return TSGPU_OK; }   // 0: tree levels of the window reduction with one thread per addition (process-wide)
    if (!strcmp(key, "h2d_overlap")) { ctx->h2d_overlap = value != 0; return TSGPU_OK; }
    if (!strcmp(key, "msm_tables")) { ctx->msm_tables = value != 0; return TSGPU_OK; }   // 0: per-window bucket sets on the plain SRS points
    if (!strcmp(key, "eval_basis")) { ctx->eval_basis = value != 0; return TSGPU_OK; }   // 0: Twist/Shout::prove interpolate and commit coefficients
    return fail(ctx, TSGPU_E_INVALID_PARAMETERS, std::string("unknown tuning key ") + key);
}

// ------------------------------------------------------------------------------------------- tables
int tsgpu_table_upload(tsgpu_ctx* ctx, const tsgpu_fr* evals, size_t n, unsigned num_vars, tsgpu_table** out) {
    if (!ctx || !out || (!evals && n)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    tsgpu_table* t = nullptr;
    int rc = table_alloc(ctx, num_vars, &t);
    if (rc) return rc;
    size_t size = (size_t)1 << num_vars;
    size_t take = n < size ? n : size;   // from_evaluations_vec pads with zeros or truncates (polynomials.rs:40-50)
    TempBuf tmp;
    TSG_CUDA(ctx, tmp.alloc(size * sizeof(fr_t), ctx->stream));
    if (take < size) TSG_CUDA(ctx, cudaMemsetAsync(tmp.as<fr_t>() + take, 0, (size - take) * sizeof(fr_t), ctx->stream));
    if (take) TSG_CUDA(ctx, cudaMemcpyAsync(tmp.p, evals, take * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    TSG_CUDA(ctx, launch_bitrev_permute(tmp.as<fr_t>(), t->d, num_vars, ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // host buffer is only borrowed for the call
    *out = t;
    return TSGPU_OK;
}

int tsgpu_table_download(tsgpu_ctx* ctx, const tsgpu_table* t, tsgpu_fr* outp) {
    if (!ctx || !t || !outp) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    size_t size = (size_t)1 << t->num_vars;
    TempBuf tmp;
    TSG_CUDA(ctx, tmp.alloc(size * sizeof(fr_t), ctx->stream));
    TSG_CUDA(ctx, launch_bitrev_permute(t->d, tmp.as<fr_t>(), t->num_vars, ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    TSG_CUDA(ctx, cudaMemcpyAsync(outp, tmp.p, size * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TSGPU_OK;
}

int tsgpu_table_clone(tsgpu_ctx* ctx, const tsgpu_table* t, tsgpu_table** out) {
    if (!ctx || !t || !out) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    tsgpu_table* c = nullptr;
    int rc = table_alloc(ctx, t->num_vars, &c);
    if (rc) return rc;
    TSG_CUDA(ctx, cudaMemcpyAsync(c->d, t->d, ((size_t)1 << t->num_vars) * sizeof(fr_t), cudaMemcpyDeviceToDevice, ctx->stream));
    *out = c;
    return TSGPU_OK;
}

unsigned tsgpu_table_num_vars(const tsgpu_table* t) { return t ? t->num_vars : 0; }

void tsgpu_table_free(tsgpu_ctx* ctx, tsgpu_table* t) {
    if (!t) return;
    if (t->d) cudaFreeAsync(t->d, ctx ? ctx->stream : nullptr);
    delete t;
}

int tsgpu_table_eq(tsgpu_ctx* ctx, const tsgpu_fr* w, unsigned num_vars, tsgpu_table** out) {
    if (!ctx || !out || (!w && num_vars)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    tsgpu_table* t = nullptr;
    int rc = table_alloc(ctx, num_vars, &t);
    if (rc) return rc;
    TempBuf wd;
    TSG_CUDA(ctx, wd.alloc((num_vars + 1) * sizeof(fr_t), ctx->stream));
    if (num_vars) TSG_CUDA(ctx, cudaMemcpyAsync(wd.p, w, num_vars * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    TSG_CUDA(ctx, launch_eq_table_bitrev(wd.as<fr_t>(), num_vars, t->d, ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = t;
    return TSGPU_OK;
}

int tsgpu_table_one_hot_rows(tsgpu_ctx* ctx, const uint64_t* idx, size_t rows, unsigned log_k, unsigned num_vars, tsgpu_table** out) {
    if (!ctx || !out || (!idx && rows)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (log_k > num_vars || rows > ((size_t)1 << (num_vars - log_k)))
        return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "one-hot rows exceed table size");
    tsgpu_table* t = nullptr;
    int rc = table_alloc(ctx, num_vars, &t);
    if (rc) return rc;
    TSG_CUDA(ctx, cudaMemsetAsync(t->d, 0, ((size_t)1 << num_vars) * sizeof(fr_t), ctx->stream));
    TempBuf id;
    TSG_CUDA(ctx, id.alloc(rows * 8, ctx->stream));
    if (rows) {
        TSG_CUDA(ctx, cudaMemcpyAsync(id.p, idx, rows * 8, cudaMemcpyHostToDevice, ctx->stream));
        TSG_CUDA(ctx, launch_one_hot_scatter(id.as<unsigned long long>(), rows, log_k, num_vars, t->d, ctx->sm_count, ctx->stream));
        ctx->launches += 1;
    }
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = t;
    return TSGPU_OK;
}

int tsgpu_table_from_u64(tsgpu_ctx* ctx, const uint64_t* v, size_t n, unsigned num_vars, tsgpu_table** out) {
    if (!ctx || !out || (!v && n)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    tsgpu_table* t = nullptr;
    int rc = table_alloc(ctx, num_vars, &t);
    if (rc) return rc;
    size_t size = (size_t)1 << num_vars;
    size_t take = n < size ? n : size;
    TSG_CUDA(ctx, cudaMemsetAsync(t->d, 0, size * sizeof(fr_t), ctx->stream));
    TempBuf src;
    TSG_CUDA(ctx, src.alloc(take * 8, ctx->stream));
    if (take) {
        TSG_CUDA(ctx, cudaMemcpyAsync(src.p, v, take * 8, cudaMemcpyHostToDevice, ctx->stream));
        TSG_CUDA(ctx, launch_fr_from_u64(src.as<unsigned long long>(), take, t->d, num_vars, 1, ctx->sm_count, ctx->stream));
        ctx->launches += 1;
    }
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = t;
    return TSGPU_OK;
}

static int read_result(tsgpu_ctx* ctx, int count, tsgpu_fr* out);
// MultilinearExtension::one_hot(num_vars, index) (src/polynomials.rs:71-82); the reference asserts index < 2^num_vars
int tsgpu_table_one_hot(tsgpu_ctx* ctx, unsigned num_vars, size_t index, tsgpu_table** out) {
    if (!ctx || !out) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (num_vars > 40 || index >= ((size_t)1 << num_vars)) {
        std::string m = "Index " + std::to_string(index) + " out of bounds for size " + std::to_string((size_t)1 << (num_vars > 40 ? 40 : num_vars));
        return fail(ctx, TSGPU_E_POLYNOMIAL, m.c_str());
    }
    const uint64_t idx = (uint64_t)index;
    return tsgpu_table_one_hot_rows(ctx, &idx, 1, num_vars, num_vars, out);
}

// MultilinearExtension::from_sparse(num_vars, &[(index, value)]) (src/polynomials.rs:52-67): zero table, then evaluations[index] = value
// in slice order - a repeated index keeps its LAST value.  The host drops the overwritten entries, the device scatters the rest.
int tsgpu_table_from_sparse(tsgpu_ctx* ctx, unsigned num_vars, const uint64_t* indices, const tsgpu_fr* values, size_t count, tsgpu_table** out) {
    if (!ctx || !out || ((!indices || !values) && count)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (num_vars > 40) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "too many variables");
    const size_t size = (size_t)1 << num_vars;
    for (size_t i = 0; i < count; ++i) if (indices[i] >= size) {
        std::string m = "Index " + std::to_string(indices[i]) + " out of bounds for size " + std::to_string(size);
        return fail(ctx, TSGPU_E_POLYNOMIAL, m.c_str());
    }
    std::vector<uint64_t> idx; std::vector<tsgpu_fr> val;
    {   // keep the last occurrence of every index
        std::vector<size_t> order(count);
        for (size_t i = 0; i < count; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return indices[a] < indices[b]; });
        for (size_t k = 0; k < count; ++k)
            if (k + 1 == count || indices[order[k + 1]] != indices[order[k]]) { idx.push_back(indices[order[k]]); val.push_back(values[order[k]]); }
    }
    tsgpu_table* t = nullptr;
    int rc = table_alloc(ctx, num_vars, &t);
    if (rc) return rc;
    TSG_CUDA(ctx, cudaMemsetAsync(t->d, 0, size * sizeof(fr_t), ctx->stream));
    TempBuf di, dv;
    TSG_CUDA(ctx, di.alloc(idx.size() * 8, ctx->stream));
    TSG_CUDA(ctx, dv.alloc(idx.size() * sizeof(fr_t), ctx->stream));
    if (!idx.empty()) {
        TSG_CUDA(ctx, cudaMemcpyAsync(di.p, idx.data(), idx.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        TSG_CUDA(ctx, cudaMemcpyAsync(dv.p, val.data(), idx.size() * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
        TSG_CUDA(ctx, launch_sparse_scatter(di.as<unsigned long long>(), dv.as<fr_t>(), idx.size(), num_vars, t->d, ctx->sm_count, ctx->stream));
        ctx->launches += 1;
    }
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = t;
    return TSGPU_OK;
}

// LessThanPolynomial::new(num_vars).to_multilinear_extension() (src/polynomials.rs:243-263): 2 * num_vars variables, generated on the device
int tsgpu_table_less_than(tsgpu_ctx* ctx, unsigned num_vars, tsgpu_table** out) {
    if (!ctx || !out) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (num_vars > 16) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "less-than table limited to 16-bit operands (2^32 entries)");
    tsgpu_table* t = nullptr;
    int rc = table_alloc(ctx, 2 * num_vars, &t);
    if (rc) return rc;
    TSG_CUDA(ctx, launch_lt_table(num_vars, t->d, ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    *out = t;
    return TSGPU_OK;
}

// MultilinearExtension::add (src/polynomials.rs:164-176); the reference asserts equal num_vars ("Number of variables must match")
int tsgpu_table_add(tsgpu_ctx* ctx, const tsgpu_table* a, const tsgpu_table* b, tsgpu_table** out) {
    if (!ctx || !a || !b || !out) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (a->num_vars != b->num_vars) return fail(ctx, TSGPU_E_POLYNOMIAL, "Number of variables must match");
    tsgpu_table* t = nullptr;
    int rc = table_alloc(ctx, a->num_vars, &t);
    if (rc) return rc;
    TSG_CUDA(ctx, launch_table_add(a->d, b->d, t->d, (size_t)1 << a->num_vars, ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    *out = t;
    return TSGPU_OK;
}
// MultilinearExtension::scalar_mul (src/polynomials.rs:179-189)
int tsgpu_table_scalar_mul(tsgpu_ctx* ctx, const tsgpu_table* a, const tsgpu_fr* scalar, tsgpu_table** out) {
    if (!ctx || !a || !scalar || !out) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    tsgpu_table* t = nullptr;
    int rc = table_alloc(ctx, a->num_vars, &t);
    if (rc) return rc;
    TSG_CUDA(ctx, launch_table_scale(a->d, t->d, (size_t)1 << a->num_vars, to_fr(scalar), ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    *out = t;
    return TSGPU_OK;
}
// MultilinearExtension::sum_evaluations (src/polynomials.rs:192-195)
int tsgpu_table_sum_evaluations(tsgpu_ctx* ctx, const tsgpu_table* t, tsgpu_fr* out) {
    if (!ctx || !t || !out) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    { KernelTimer kt(ctx, "table_sum");
      TSG_CUDA(ctx, launch_table_sum(t->d, (size_t)1 << t->num_vars, ctx->partials, ctx->ticket, ctx->dev_out, ctx->sm_count, ctx->stream)); }
    ctx->launches += 1;
    return read_result(ctx, 1, out);
}

// ---- lookup-argument building blocks (csrc/lookup.cu): the one-hot matrix ra(x, j) = [idx_j == x] applied without materialising it
// out[x] = sum over j < n with idx[j] == x of weights[j], x < 2^log_k   (weights: a table of at least n entries)
int tsgpu_table_scatter_add(tsgpu_ctx* ctx, const tsgpu_table* weights, const uint64_t* idx, size_t n, unsigned log_k, tsgpu_table** out) {
    if (!ctx || !weights || !out || (!idx && n)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (log_k > 32 || n > ((size_t)1 << weights->num_vars)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "more indices than weights");
    const size_t K = (size_t)1 << log_k;
    for (size_t j = 0; j < n; ++j) if (idx[j] >= K) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Lookup index out of bounds");   // shout.rs:44-48
    tsgpu_table* t = nullptr;
    int rc = table_alloc(ctx, log_k, &t);
    if (rc) return rc;
    TempBuf di, acc;
    TSG_CUDA(ctx, di.alloc(n * 8, ctx->stream));
    TSG_CUDA(ctx, acc.alloc(K * 64, ctx->stream));
    if (n) TSG_CUDA(ctx, cudaMemcpyAsync(di.p, idx, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    TSG_CUDA(ctx, cudaMemsetAsync(acc.p, 0, K * 64, ctx->stream));
    TSG_CUDA(ctx, launch_weighted_hist(weights->d, weights->num_vars, di.as<unsigned long long>(), n, acc.as<unsigned long long>(), ctx->sm_count, ctx->stream));
    TSG_CUDA(ctx, launch_limb_sums_to_table(acc.as<unsigned long long>(), log_k, t->d, ctx->sm_count, ctx->stream));
    ctx->launches += n ? 2 : 1;
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // `idx` is borrowed
    *out = t;
    return TSGPU_OK;
}
// out[j] = src[idx[j]] for j < n, zero for n <= j < 2^num_vars
int tsgpu_table_gather(tsgpu_ctx* ctx, const tsgpu_table* src, const uint64_t* idx, size_t n, unsigned num_vars, tsgpu_table** out) {
    if (!ctx || !src || !out || (!idx && n)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (num_vars > 40 || n > ((size_t)1 << num_vars)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "more indices than output entries");
    const size_t K = (size_t)1 << src->num_vars;
    for (size_t j = 0; j < n; ++j) if (idx[j] >= K) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Lookup index out of bounds");
    tsgpu_table* t = nullptr;
    int rc = table_alloc(ctx, num_vars, &t);
    if (rc) return rc;
    TempBuf di;
    TSG_CUDA(ctx, di.alloc(n * 8, ctx->stream));
    if (n) TSG_CUDA(ctx, cudaMemcpyAsync(di.p, idx, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    TSG_CUDA(ctx, launch_table_gather(src->d, src->num_vars, di.as<unsigned long long>(), n, num_vars, t->d, ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = t;
    return TSGPU_OK;
}
// field_utils::inner_product (src/utils.rs:210-213) of two tables of equal size: lazy 512-bit dot product, one pass
int tsgpu_table_inner_product(tsgpu_ctx* ctx, const tsgpu_table* a, const tsgpu_table* b, tsgpu_fr* out) {
    if (!ctx || !a || !b || !out) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (a->num_vars != b->num_vars) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Vector lengths must match");
    TSG_CUDA(ctx, launch_dot(a->d, b->d, (size_t)1 << a->num_vars, ctx->partials, ctx->ticket, ctx->dev_out, ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    return read_result(ctx, 1, out);
}

// ---- read/write memory (Twist) tables over (cell x, cycle j), reference index x + 2^log_cells * j (csrc/lookup.cu)
// Val(x, j): content of cell x just before operation j of the trace (zero-initialised memory, MemoryTrace semantics, src/twist.rs:48-70)
int tsgpu_table_memory_values(tsgpu_ctx* ctx, const uint64_t* addresses, const uint8_t* is_write, const tsgpu_fr* values, size_t n, unsigned log_cells,
                              unsigned log_cycles, tsgpu_table** out) {
    if (!ctx || !out || ((!addresses || !is_write || !values) && n)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (log_cells + log_cycles > 30 || n > ((size_t)1 << log_cycles)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "memory table too large");
    for (size_t j = 0; j < n; ++j) if (addresses[j] >> log_cells) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Address out of bounds");   // twist.rs:49-53
    tsgpu_table* t = nullptr;
    int rc = table_alloc(ctx, log_cells + log_cycles, &t);
    if (rc) return rc;
    // the writes, each with the next write to the same address (or T): one backward pass over the trace
    const uint64_t T = (uint64_t)1 << log_cycles;
    std::vector<uint64_t> wa, wj, wnext; std::vector<tsgpu_fr> wv;
    {
        std::vector<uint64_t> next_of_cell((size_t)1 << log_cells, T);
        size_t nw = 0;
        for (size_t j = 0; j < n; ++j) nw += is_write[j] ? 1 : 0;
        wa.resize(nw); wj.resize(nw); wnext.resize(nw); wv.resize(nw);
        size_t w = nw;
        for (size_t j = n; j-- > 0;) if (is_write[j]) {
            --w; wa[w] = addresses[j]; wj[w] = j; wnext[w] = next_of_cell[addresses[j]]; wv[w] = values[j];
            next_of_cell[addresses[j]] = j + 1;    // the run of the previous write to this cell ends where this write takes effect (cycle j + 1)
        }
    }
    const size_t nw = wa.size();
    TSG_CUDA(ctx, cudaMemsetAsync(t->d, 0, ((size_t)1 << (log_cells + log_cycles)) * sizeof(fr_t), ctx->stream));
    TempBuf da, dj, dn, dv;
    TSG_CUDA(ctx, da.alloc(nw * 8, ctx->stream)); TSG_CUDA(ctx, dj.alloc(nw * 8, ctx->stream)); TSG_CUDA(ctx, dn.alloc(nw * 8, ctx->stream));
    TSG_CUDA(ctx, dv.alloc(nw * sizeof(fr_t), ctx->stream));
    if (nw) {
        TSG_CUDA(ctx, cudaMemcpyAsync(da.p, wa.data(), nw * 8, cudaMemcpyHostToDevice, ctx->stream));
        TSG_CUDA(ctx, cudaMemcpyAsync(dj.p, wj.data(), nw * 8, cudaMemcpyHostToDevice, ctx->stream));
        TSG_CUDA(ctx, cudaMemcpyAsync(dn.p, wnext.data(), nw * 8, cudaMemcpyHostToDevice, ctx->stream));
        TSG_CUDA(ctx, cudaMemcpyAsync(dv.p, wv.data(), nw * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
        TSG_CUDA(ctx, launch_val_fill(da.as<unsigned long long>(), dj.as<unsigned long long>(), dn.as<unsigned long long>(), dv.as<fr_t>(), nw, log_cells, log_cycles,
                                      t->d, ctx->sm_count, ctx->stream));
        ctx->launches += 1;
    }
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = t;
    return TSGPU_OK;
}
// out[addresses[j] + 2^log_cells * j] = weights[j] for the j < n with select[j] == flag, zero elsewhere: the address one-hot matrix with weighted rows
int tsgpu_table_one_hot_weighted(tsgpu_ctx* ctx, const tsgpu_table* weights, const uint64_t* addresses, const uint8_t* select, int flag, size_t n,
                                 unsigned log_cells, tsgpu_table** out) {
    if (!ctx || !weights || !out || ((!addresses || !select) && n)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    const unsigned t = weights->num_vars;
    if (log_cells + t > 30 || n > ((size_t)1 << t)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "memory table too large");
    for (size_t j = 0; j < n; ++j) if (addresses[j] >> log_cells) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Address out of bounds");
    tsgpu_table* o = nullptr;
    int rc = table_alloc(ctx, log_cells + t, &o);
    if (rc) return rc;
    TSG_CUDA(ctx, cudaMemsetAsync(o->d, 0, ((size_t)1 << (log_cells + t)) * sizeof(fr_t), ctx->stream));
    TempBuf da, ds;
    TSG_CUDA(ctx, da.alloc(n * 8, ctx->stream)); TSG_CUDA(ctx, ds.alloc(n, ctx->stream));
    if (n) {
        TSG_CUDA(ctx, cudaMemcpyAsync(da.p, addresses, n * 8, cudaMemcpyHostToDevice, ctx->stream));
        TSG_CUDA(ctx, cudaMemcpyAsync(ds.p, select, n, cudaMemcpyHostToDevice, ctx->stream));
        TSG_CUDA(ctx, launch_one_hot_weighted(weights->d, da.as<unsigned long long>(), ds.as<unsigned char>(), (unsigned char)(flag != 0), n, log_cells, t, o->d,
                                              ctx->sm_count, ctx->stream));
        ctx->launches += 1;
    }
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = o;
    return TSGPU_OK;
}
// table(x, j) <- rows(j) - table(x, j) in place: `rows` has log_cycles variables, `table` log_cells + log_cycles (reference index x + 2^log_cells j)
int tsgpu_table_broadcast_rows_minus(tsgpu_ctx* ctx, const tsgpu_table* rows, tsgpu_table* table) {
    if (!ctx || !rows || !table) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (rows->num_vars > table->num_vars) return fail(ctx, TSGPU_E_POLYNOMIAL, "Number of variables must match");
    TSG_CUDA(ctx, launch_broadcast_rows_minus(rows->d, rows->num_vars, table->d, (size_t)1 << table->num_vars, ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    return TSGPU_OK;
}
// elementwise product of two tables of equal size
int tsgpu_table_mul(tsgpu_ctx* ctx, const tsgpu_table* a, const tsgpu_table* b, tsgpu_table** out) {
    if (!ctx || !a || !b || !out) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (a->num_vars != b->num_vars) return fail(ctx, TSGPU_E_POLYNOMIAL, "Number of variables must match");
    tsgpu_table* t = nullptr;
    int rc = table_alloc(ctx, a->num_vars, &t);
    if (rc) return rc;
    TSG_CUDA(ctx, launch_table_mul(a->d, b->d, t->d, (size_t)1 << a->num_vars, ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    *out = t;
    return TSGPU_OK;
}
// out[a] = LT~(a, point), a in {0,1}^num_vars: [a < c] in the natural integer order (bit num_vars - 1 most significant), multilinear in c, at c = point
int tsgpu_table_lt_point(tsgpu_ctx* ctx, const tsgpu_fr* point, unsigned num_vars, tsgpu_table** out) {
    if (!ctx || !out || (!point && num_vars)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    tsgpu_table* t = nullptr;
    int rc = table_alloc(ctx, num_vars, &t);
    if (rc) return rc;
    TempBuf pd;
    TSG_CUDA(ctx, pd.alloc((num_vars + 1) * sizeof(fr_t), ctx->stream));
    if (num_vars) TSG_CUDA(ctx, cudaMemcpyAsync(pd.p, point, num_vars * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    TSG_CUDA(ctx, launch_lt_point_table(pd.as<fr_t>(), num_vars, t->d, ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = t;
    return TSGPU_OK;
}

// ------------------------------------------------------------------------------- evaluate / partial
// results that a kernel wrote directly into the pinned mirror: wait for the stream, copy out
static int read_host_result(tsgpu_ctx* ctx, int count, tsgpu_fr* out) {
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(out, ctx->host_out, count * sizeof(fr_t));
    return TSGPU_OK;
}
static int read_result(tsgpu_ctx* ctx, int count, tsgpu_fr* out) {
    TSG_CUDA(ctx, cudaMemcpyAsync(ctx->host_out, ctx->dev_out, count * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(out, ctx->host_out, count * sizeof(fr_t));
    return TSGPU_OK;
}

int tsgpu_table_partial_evaluate(tsgpu_ctx* ctx, const tsgpu_table* t, const tsgpu_fr* fixed, unsigned k, tsgpu_table** out) {
    if (!ctx || !t || !out || (!fixed && k)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (k > t->num_vars) return fail(ctx, TSGPU_E_POLYNOMIAL, "Cannot fix more variables than available");   // polynomials.rs:128
    if (k == 0) return tsgpu_table_clone(ctx, t, out);   // polynomials.rs:130-132
    tsgpu_table* o = nullptr;
    int rc = table_alloc(ctx, t->num_vars - k, &o);
    if (rc) return rc;
    size_t rows = (size_t)1 << k, cols = (size_t)1 << (t->num_vars - k);
    TempBuf fx, W, part;
    TSG_CUDA(ctx, fx.alloc(k * sizeof(fr_t), ctx->stream));
    TSG_CUDA(ctx, W.alloc(rows * sizeof(fr_t), ctx->stream));
    TSG_CUDA(ctx, cudaMemcpyAsync(fx.p, fixed, k * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    TSG_CUDA(ctx, launch_eq_table_bitrev(fx.as<fr_t>(), k, W.as<fr_t>(), ctx->sm_count, ctx->stream));
    size_t nsplit = colsum_splits(rows, cols, ctx->sm_count);
    TSG_CUDA(ctx, part.alloc(nsplit > 1 ? nsplit * cols * sizeof(fr_t) : 32, ctx->stream));
    TSG_CUDA(ctx, launch_colsum(t->d, W.as<fr_t>(), rows, cols, nsplit, part.as<fr_t>(), o->d, ctx->stream));
    ctx->launches += 2 + (nsplit > 1 ? 1 : 0);
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // `fixed` is borrowed
    *out = o;
    return TSGPU_OK;
}

int tsgpu_table_evaluate(tsgpu_ctx* ctx, const tsgpu_table* t, const tsgpu_fr* point, tsgpu_fr* out) {
    if (!ctx || !t || !out || (!point && t->num_vars)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    const unsigned nv = t->num_vars;
    const unsigned l = nv < 16 ? nv : 16, h = nv - l;
    size_t rows = (size_t)1 << h, cols = (size_t)1 << l;
    TempBuf pt, Whi, Wlo, colv, part;
    TSG_CUDA(ctx, pt.alloc((nv + 1) * sizeof(fr_t), ctx->stream));
    if (nv) TSG_CUDA(ctx, cudaMemcpyAsync(pt.p, point, nv * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    TSG_CUDA(ctx, Wlo.alloc(cols * sizeof(fr_t), ctx->stream));
    TSG_CUDA(ctx, launch_eq_table_bitrev(pt.as<fr_t>() + h, l, Wlo.as<fr_t>(), ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    const fr_t* vec = t->d;
    if (h) {
        TSG_CUDA(ctx, Whi.alloc(rows * sizeof(fr_t), ctx->stream));
        TSG_CUDA(ctx, colv.alloc(cols * sizeof(fr_t), ctx->stream));
        TSG_CUDA(ctx, launch_eq_table_bitrev(pt.as<fr_t>(), h, Whi.as<fr_t>(), ctx->sm_count, ctx->stream));
        size_t nsplit = colsum_splits(rows, cols, ctx->sm_count);
        TSG_CUDA(ctx, part.alloc(nsplit > 1 ? nsplit * cols * sizeof(fr_t) : 32, ctx->stream));
        TSG_CUDA(ctx, launch_colsum(t->d, Whi.as<fr_t>(), rows, cols, nsplit, part.as<fr_t>(), colv.as<fr_t>(), ctx->stream));
        ctx->launches += 2 + (nsplit > 1 ? 1 : 0);
        vec = colv.as<fr_t>();
    }
    TSG_CUDA(ctx, launch_dot(vec, Wlo.as<fr_t>(), cols, ctx->partials, ctx->ticket, ctx->dev_out, ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    return read_result(ctx, 1, out);
}

int tsgpu_table_bind(tsgpu_ctx* ctx, tsgpu_table* t, const tsgpu_fr* r) {
    if (!ctx || !t || !r) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (t->num_vars == 0) return fail(ctx, TSGPU_E_POLYNOMIAL, "Cannot fix more variables than available");
    { KernelTimer kt(ctx, "bind");
      TSG_CUDA(ctx, launch_bind(t->d, (size_t)1 << t->num_vars, to_fr(r), ctx->sm_count, ctx->stream)); }
    ctx->launches += 1;
    t->num_vars -= 1;
    return TSGPU_OK;
}

int tsgpu_mle_evaluate(tsgpu_ctx* ctx, const tsgpu_fr* evals, unsigned num_vars, const tsgpu_fr* point, tsgpu_fr* out) {
    tsgpu_table* t = nullptr;
    int rc = tsgpu_table_upload(ctx, evals, (size_t)1 << num_vars, num_vars, &t);
    if (rc) return rc;
    rc = tsgpu_table_evaluate(ctx, t, point, out);
    tsgpu_table_free(ctx, t);
    return rc;
}

int tsgpu_mle_partial_evaluate(tsgpu_ctx* ctx, const tsgpu_fr* evals, unsigned num_vars, const tsgpu_fr* fixed, unsigned k, tsgpu_fr* out) {
    if (k > num_vars) return fail(ctx, TSGPU_E_POLYNOMIAL, "Cannot fix more variables than available");
    tsgpu_table *t = nullptr, *o = nullptr;
    int rc = tsgpu_table_upload(ctx, evals, (size_t)1 << num_vars, num_vars, &t);
    if (rc) return rc;
    rc = tsgpu_table_partial_evaluate(ctx, t, fixed, k, &o);
    if (!rc) rc = tsgpu_table_download(ctx, o, out);
    tsgpu_table_free(ctx, t);
    tsgpu_table_free(ctx, o);
    return rc;
}

// ------------------------------------------------------------------------------------- sum-check
static int sc_tail_leave(tsgpu_sc* sc);
int tsgpu_sc_begin(tsgpu_ctx* ctx, tsgpu_table* const* tables, int d, tsgpu_sc** out) {
    if (!ctx || !tables || !out) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (d < 1 || d > SC_MAX_TABLES) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "sum-check supports products of 1..3 tables");
    for (int i = 0; i < d; ++i) {
        if (!tables[i]) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null table");
        if (tables[i]->num_vars != tables[0]->num_vars) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "Number of variables must match");
    }
    tsgpu_sc* sc = new (std::nothrow) tsgpu_sc;
    if (!sc) return fail(ctx, TSGPU_E_PROOF_GENERATION, "out of host memory");
    sc->ctx = ctx; sc->d = d; sc->vars_left = tables[0]->num_vars;
    for (int i = 0; i < d; ++i) sc->tables[i] = tables[i];
    *out = sc;
    return TSGPU_OK;
}

unsigned tsgpu_sc_num_vars(const tsgpu_sc* sc) { return sc ? sc->vars_left : 0; }
int tsgpu_sc_exclusive(tsgpu_sc* sc, int on) {
    if (!sc) return TSGPU_E_INVALID_PARAMETERS;
    if (!on) { int rc = sc_tail_leave(sc); if (rc) return rc; }
    sc->exclusive = on != 0;
    return TSGPU_OK;
}

static ScTables sc_tabs(const tsgpu_sc* sc) {
    ScTables t; for (int i = 0; i < SC_MAX_TABLES; ++i) t.t[i] = i < sc->d ? sc->tables[i]->d : nullptr;
    return t;
}

// ---- persistent tail (sumcheck.cu): entered by tsgpu_sc_bind_eval_claim once the tables fit one CTA's shared memory; every other entry point first
// makes the resident kernel write the tables back (sc_tail_leave), so callers may mix the calls freely
static int sc_tail_fail(tsgpu_sc* sc) {
    sc_tail_post_command(sc->ctx->tail_box, SC_TAIL_ABORT);
    cudaStreamSynchronize(sc->ctx->stream);
    sc->tail_active = false;
    return fail(sc->ctx, TSGPU_E_PROOF_GENERATION, "sum-check tail kernel did not answer");
}
static int sc_tail_leave(tsgpu_sc* sc) {
    if (!sc->tail_active) return TSGPU_OK;
    sc_tail_post_command(sc->ctx->tail_box, SC_TAIL_FLUSH);
    sc->tail_active = false;
    TSG_CUDA(sc->ctx, cudaStreamSynchronize(sc->ctx->stream));
    return TSGPU_OK;
}
// one step of the tail: fold with r, return g(0), g(2) of the next round (or, with no variable left after the fold, the two bound table values)
static int sc_tail_step(tsgpu_sc* sc, const fr_t& r, fr_t v[2]) {
    tsgpu_ctx* ctx = sc->ctx;
    if (!sc->tail_active) {
        TSG_CUDA(ctx, launch_sc_tail(sc->tables[0]->d, sc->tables[1]->d, (size_t)1 << sc->vars_left, r, ctx->tail_box, ctx->stream));
        ctx->launches += 1;
        sc->tail_active = true; sc->tail_seq = 0;
    } else {
        sc_tail_post_challenge(ctx->tail_box, r, sc->tail_seq);
    }
    sc->tail_seq += 1;
    if (!sc_tail_wait(ctx->tail_box, sc->tail_seq, ctx->stream, v)) return sc_tail_fail(sc);
    for (int i = 0; i < sc->d; ++i) sc->tables[i]->num_vars -= 1;
    sc->vars_left -= 1;
    if (sc->vars_left == 0) {                     // the kernel has written the bound values to the tables and is leaving
        sc->tail_active = false;
        TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return TSGPU_OK;
}

int tsgpu_sc_round_eval(tsgpu_sc* sc, tsgpu_fr evals[4]) {
    if (!sc || !evals) return TSGPU_E_INVALID_PARAMETERS;
    tsgpu_ctx* ctx = sc->ctx;
    if (int lrc = sc_tail_leave(sc)) return lrc;
    if (sc->vars_left == 0) return fail(ctx, TSGPU_E_SUMCHECK, "no variables left to evaluate");
    KernelTimer kt(ctx, "sc_round_eval");
    // the finishing block writes the four values straight into the pinned host mirror (unified addressing: no D2H copy to enqueue per round)
    TSG_CUDA(ctx, launch_round_eval(sc->d, sc_tabs(sc), (size_t)1 << sc->vars_left, ctx->partials, ctx->ticket, ctx->host_out, ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    return read_host_result(ctx, 4, evals);
}

// tsgpu_sc_round_eval for d = 2 when the caller vouches for g(0) + g(1) = claim (it must then check the claim by other means: the host loop of
// host/sumcheck_host.cpp does so at the end of the protocol); for d != 2 the claim is ignored and the full evaluation runs
int tsgpu_sc_round_eval_claim(tsgpu_sc* sc, const tsgpu_fr* claim, tsgpu_fr evals[4]) {
    if (!sc || !evals || !claim) return TSGPU_E_INVALID_PARAMETERS;
    tsgpu_ctx* ctx = sc->ctx;
    if (sc->vars_left == 0) return fail(ctx, TSGPU_E_SUMCHECK, "no variables left to evaluate");
    if (int lrc = sc_tail_leave(sc)) return lrc;
    KernelTimer kt(ctx, "sc_round_eval");
    const fr_t cl = to_fr(claim);
    TSG_CUDA(ctx, launch_round_eval(sc->d, sc_tabs(sc), (size_t)1 << sc->vars_left, ctx->partials, ctx->ticket, ctx->host_out, ctx->sm_count, ctx->stream,
                                    sc->d == 2 ? &cl : nullptr));
    ctx->launches += 1;
    return read_host_result(ctx, 4, evals);
}

int tsgpu_sc_bind(tsgpu_sc* sc, const tsgpu_fr* r) {
    if (!sc || !r) return TSGPU_E_INVALID_PARAMETERS;
    tsgpu_ctx* ctx = sc->ctx;
    if (sc->vars_left == 0) return fail(ctx, TSGPU_E_SUMCHECK, "no variables left to bind");
    if (sc->tail_active && sc->vars_left == 1) { fr_t fin[2]; return sc_tail_step(sc, to_fr(r), fin); }   // last fold inside the resident kernel
    if (int lrc = sc_tail_leave(sc)) return lrc;
    for (int i = 0; i < sc->d; ++i) {
        TSG_CUDA(ctx, launch_bind(sc->tables[i]->d, (size_t)1 << sc->vars_left, to_fr(r), ctx->sm_count, ctx->stream));
        ctx->launches += 1;
        sc->tables[i]->num_vars -= 1;
    }
    sc->vars_left -= 1;
    return TSGPU_OK;
}

static int sc_bind_eval_impl(tsgpu_sc* sc, const tsgpu_fr* r, const tsgpu_fr* claim, tsgpu_fr evals[4]) {
    if (!sc || !r || !evals) return TSGPU_E_INVALID_PARAMETERS;
    tsgpu_ctx* ctx = sc->ctx;
    if (sc->vars_left < 2) return fail(ctx, TSGPU_E_SUMCHECK, "bind_eval needs at least two unbound variables");
    if (claim && sc->d == 2 && sc->exclusive && ctx->sc_tail && ctx->tail_box && (sc->tail_active || sc->vars_left - 1 <= SC_TAIL_MAX_LOG)) {
        fr_t v[2];
        int trc = sc_tail_step(sc, to_fr(r), v);
        if (trc) return trc;
        // the four values from g(0), g(2) and the claim, as the per-round kernel's epilogue computes them
        const fr_t g1 = to_fr(claim) - v[0], dd = v[1] - g1;
        const fr_t ev[4] = {v[0], g1, v[1], v[0] + dd + dd + dd};
        memcpy(evals, ev, sizeof(ev));
        return TSGPU_OK;
    }
    if (int lrc = sc_tail_leave(sc)) return lrc;
    KernelTimer kt(ctx, "sc_bind_eval");
    fr_t cl; if (claim) cl = to_fr(claim);
    TSG_CUDA(ctx, launch_bind_eval(sc->d, sc_tabs(sc), (size_t)1 << sc->vars_left, to_fr(r), claim ? &cl : nullptr, ctx->partials, ctx->ticket, ctx->host_out,
                                   ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    for (int i = 0; i < sc->d; ++i) sc->tables[i]->num_vars -= 1;
    sc->vars_left -= 1;
    return read_host_result(ctx, 4, evals);
}
int tsgpu_sc_bind_eval(tsgpu_sc* sc, const tsgpu_fr* r, tsgpu_fr evals[4]) { return sc_bind_eval_impl(sc, r, nullptr, evals); }
// same, given the claim of the round being evaluated (g_k(r) of the round just bound): g(0) + g(1) = claim is then an identity
// (the reference checks it every round, src/sumcheck.rs:77-84; it can only fail in round 0, which tsgpu_sc_round_eval computes in full),
// so the kernel sums g(0) and g(2) only and returns g(1) = claim - g(0).  Same four values.
int tsgpu_sc_bind_eval_claim(tsgpu_sc* sc, const tsgpu_fr* r, const tsgpu_fr* claim, tsgpu_fr evals[4]) {
    if (!claim) return TSGPU_E_INVALID_PARAMETERS;
    return sc_bind_eval_impl(sc, r, claim, evals);
}

int tsgpu_sc_final(tsgpu_sc* sc, tsgpu_fr* finals) {
    if (!sc || !finals) return TSGPU_E_INVALID_PARAMETERS;
    tsgpu_ctx* ctx = sc->ctx;
    if (sc->vars_left != 0) return fail(ctx, TSGPU_E_SUMCHECK, "variables remain unbound");
    if (int lrc = sc_tail_leave(sc)) return lrc;
    for (int i = 0; i < sc->d; ++i)
        TSG_CUDA(ctx, cudaMemcpyAsync(ctx->dev_out + i, sc->tables[i]->d, sizeof(fr_t), cudaMemcpyDeviceToDevice, ctx->stream));
    return read_result(ctx, sc->d, finals);
}

void tsgpu_sc_end(tsgpu_sc* sc) {
    if (sc && sc->tail_active) {                  // abandoned mid-protocol (an error path of the caller): the resident kernel must not outlive the handle
        sc_tail_post_command(sc->ctx->tail_box, SC_TAIL_ABORT);
        cudaStreamSynchronize(sc->ctx->stream);
    }
    delete sc;
}

}  // extern "C"
