// api_kzg.cu - C ABI (include/tsgpu.h): SRS handles, G1 MSM, KZGCommitment::commit / open.
#include <cstring>
#include <new>
#include <vector>
#include "context.cuh"
#include "msm.cuh"
#include "poly.cuh"
#include "interp.cuh"
#include "mle.cuh"
#include "lagrange.cuh"
#include <map>
#include <mutex>
#include <tuple>
#include "../host/field64.hpp"

using namespace tsg;
using tsg::host::Fq64;
using tsg::host::Fr64;
using tsg::host::G1J;

struct tsgpu_srs {
    g1_affine* d = nullptr;   // n affine points, identity = (0,0)
    size_t n = 0;
    // evaluation-basis companion (lagrange.cu): for a domain size m (power of two) the points [L_j(tau)]_1, j < m, of the
    // Lagrange basis on the nodes 0..m-1.  Built from the trapdoor, which setup_params holds (src/utils.rs:84,107);
    // an SRS uploaded as bare points has none and the coefficient path (interpolate, then commit) is used.
    bool has_tau = false;
    tsgpu_fr tau;
    std::map<size_t, g1_affine*> lagrange;
    // precomputed window tables (msm.cu): table[w * n + i] = 2^(c w) * point_i, so that all digit positions of a scalar feed
    // ONE bucket set: fewer, wider windows (c = 20 at 2^20 points: 13 additions per point instead of 16) and a single reduction
    g1_affine* table = nullptr; unsigned table_c = 0;
    std::map<size_t, std::pair<g1_affine*, unsigned>> lagrange_table;
    // second, short table per Lagrange basis for value vectors whose entries fit 64 bits (addresses, u64 memory values, table indices):
    // 16-bit windows, only the five lowest - the bucket set is 16 times smaller, and so is the window reduction
    std::map<size_t, g1_affine*> lagrange_short_table;
    // per-rank slices of an evaluation basis (sharded proving): nodes first .. first + count - 1 of the m-node domain, with their tables
    struct Slice { g1_affine* pts = nullptr; g1_affine* table = nullptr; unsigned table_c = 0; g1_affine* short_table = nullptr; };
    std::map<std::tuple<size_t, size_t, size_t>, Slice> lagrange_slices;
    // the lazy caches above are filled through const handles (tsgpu_srs_lagrange_prepare*): one lock per handle serialises the fills, so a
    // params handle may be shared by proofs running on several host threads (each with its own context / stream)
    mutable std::mutex cache_mu;
};
// a base array the MSM can run on: plain points and, optionally, their window tables
struct MsmBasis { const g1_affine* pts; size_t n; const g1_affine* table; unsigned table_c; const g1_affine* short_table = nullptr; unsigned short_c = 0, short_windows = 0; };
struct tsgpu_poly {
    fr_t* d = nullptr;        // n coefficients, low -> high, natural order
    size_t n = 0;
    bool short64 = false;          // every entry is known to fit 64 bits (built by tsgpu_poly_from_u64 and not modified since): the commit pass skips its probe
    cudaEvent_t ready = nullptr;   // set while a side-stream upload is in flight (tsgpu_poly_upload_padded_async): tsgpu_poly_wait orders the context's stream behind it
};

static_assert(sizeof(g1_affine) == 64 && sizeof(g1_jac) == 96 && sizeof(g1_xyzz) == 128, "point layouts");
static_assert(sizeof(tsgpu_g1) == 96 && sizeof(tsgpu_g1a) == 64, "ABI point layouts");

namespace {

// combine W window sums S_w (Jacobian, low window first): sum_w 2^(c w) S_w
G1J combine_windows(const g1_jac* win, unsigned W, unsigned c) {
    G1J acc = G1J::identity();
    for (unsigned w = W; w-- > 0;) {
        for (unsigned k = 0; k < c; ++k) acc = acc.dbl();
        G1J s; memcpy(&s, &win[w], 96);
        acc = acc.add(s);
    }
    return acc;
}

// K independent device MSMs in one pass (bases and scalars resident); results to the host as Jacobian points.
// Mode: precomputed tables (one shared bucket set per job) when every job has a table of the same window width and uses at
// least a quarter of it; otherwise per-window bucket sets on the plain points.
int msm_device_batch(tsgpu_ctx* ctx, int K, const MsmBasis* basis, const fr_t* const* scalars, const size_t* n, tsgpu_g1* out, bool maybe_short = false,
                     bool known_short = false) {
    size_t nmax = 0;
    for (int k = 0; k < K; ++k) nmax = n[k] > nmax ? n[k] : nmax;
    if (nmax == 0) { G1J id = G1J::identity(); for (int k = 0; k < K; ++k) memcpy(&out[k], &id, 96); return TSGPU_OK; }
    bool shared = ctx->msm_tables;
    for (int k = 0; k < K && shared; ++k)
        shared = basis[k].table && basis[k].table_c == basis[0].table_c && n[k] * 4 >= basis[k].n && (255 + basis[k].table_c - 1) / basis[k].table_c * basis[k].n < ((size_t)1 << 31);
    // value vectors (commitments over the Lagrange basis) are usually short scalars: probe, and if every entry fits 64 bits use the short tables
    bool use_short = false;
    if (maybe_short && shared) {
        use_short = true;
        for (int k = 0; k < K; ++k) use_short = use_short && basis[k].short_table && basis[k].short_c == basis[0].short_c;
        if (use_short && !known_short) {
            unsigned* flag = (unsigned*)(ctx->dev_out + 7);   // last result slot doubles as the probe flag
            TSG_CUDA(ctx, cudaMemsetAsync(flag, 0, 4, ctx->stream));
            for (int k = 0; k < K; ++k) TSG_CUDA(ctx, launch_scalar_probe(scalars[k], n[k], flag, ctx->sm_count, ctx->stream));
            ctx->launches += K;
            unsigned* hflag = (unsigned*)(ctx->host_out + 7);
            TSG_CUDA(ctx, cudaMemcpyAsync(hflag, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
            TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            use_short = *hflag == 0;
        }
    }
    const unsigned c = use_short ? basis[0].short_c : shared ? basis[0].table_c : msm_window_bits(nmax);
    MsmLayout L;
    std::vector<g1_jac> raw;
    unsigned counts[2] = {0, 0};   // work items, bucket entries
    unsigned char* scratch_p = nullptr;
    {
        size_t bytes = msm_scratch_bytes(nmax, K, c, shared, &L, use_short ? basis[0].short_windows : 0);
        cudaError_t aerr;
        scratch_p = (unsigned char*)arena_get(ctx, tsgpu_ctx::ARENA_MSM, bytes, &aerr);
        if (!scratch_p) return cuda_fail(ctx, aerr, "cudaMalloc(msm scratch)");
        MsmJob jobs[MSM_MAX_BATCH];
        for (int k = 0; k < K; ++k) jobs[k] = MsmJob{use_short ? basis[k].short_table : shared ? basis[k].table : basis[k].pts, basis[k].n, scalars[k], n[k]};
        unsigned launches = 0;
        cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        if (ctx->timing) for (auto& x : ev) cudaEventCreate(&x);
        {
            KernelTimer kt(ctx, "msm_total");
            TSG_CUDA(ctx, msm_run(jobs, K, L, scratch_p, ctx->sm_count, ctx->stream, &launches, ctx->timing ? ev : nullptr));
        }
        if (ctx->timing) {
            // four phases share five events (timers_collect destroys each event once)
            static const char* names[4] = {"msm_sort", "msm_accumulate", "msm_merge", "msm_reduce"};
            for (int k = 0; k < 4; ++k) ctx->pending.push_back({names[k], ev[k], ev[k + 1]});
        }
        ctx->launches += launches;
        raw.resize((size_t)L.sets * (L.span_bits + 2));
        const size_t raw_bytes = raw.size() * sizeof(g1_jac);
        if (raw_bytes + sizeof(counts) <= tsgpu_ctx::HOST_MSM_BYTES) {        // through the pinned staging block: two plain DMA transfers
            TSG_CUDA(ctx, cudaMemcpyAsync(ctx->host_msm, scratch_p + L.window_out, raw_bytes, cudaMemcpyDeviceToHost, ctx->stream));
            TSG_CUDA(ctx, cudaMemcpyAsync(ctx->host_msm + raw_bytes, scratch_p + L.n_items, sizeof(counts), cudaMemcpyDeviceToHost, ctx->stream));
            TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            memcpy(raw.data(), ctx->host_msm, raw_bytes); memcpy(counts, ctx->host_msm + raw_bytes, sizeof(counts));
        } else {
            TSG_CUDA(ctx, cudaMemcpyAsync(raw.data(), scratch_p + L.window_out, raw_bytes, cudaMemcpyDeviceToHost, ctx->stream));
            TSG_CUDA(ctx, cudaMemcpyAsync(counts, scratch_p + L.n_items, sizeof(counts), cudaMemcpyDeviceToHost, ctx->stream));
            TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        timers_collect(ctx);
    }
    for (int k = 0; k < K; ++k) ctx->msm_points += n[k];
    const unsigned per_set = L.span_bits + 2;
    std::vector<g1_jac> win(L.sets);
    ctx->msm_entries += counts[1];
    ctx->msm_calls += 1;
    // bucket-set sums: S = span * sum_k 2^k P[k] + P[span_bits] + P[span_bits + 1]  (Horner over the index-bit sums, then the span-local part)
    unsigned logS = 0; while ((1u << logS) < L.span) ++logS;
    for (unsigned w = 0; w < L.sets; ++w) {
        const g1_jac* P = raw.data() + (size_t)w * per_set;
        G1J acc = G1J::identity();
        for (unsigned k = L.span_bits; k-- > 0;) { acc = acc.dbl(); G1J t; memcpy(&t, &P[k], 96); acc = acc.add(t); }
        for (unsigned d = 0; d < logS; ++d) acc = acc.dbl();
        for (unsigned h = 0; h < 2; ++h) { G1J l; memcpy(&l, &P[L.span_bits + h], 96); acc = acc.add(l); }
        memcpy(&win[w], &acc, 96);
    }
    const unsigned per_job = L.sets / L.K;
    for (int k = 0; k < K; ++k) {
        G1J r = combine_windows(win.data() + (size_t)k * per_job, per_job, L.c);   // one set per job in table mode: nothing to combine
        memcpy(&out[k], &r, 96);
    }
    return TSGPU_OK;
}
int msm_device(tsgpu_ctx* ctx, const MsmBasis& basis, const fr_t* scalars, size_t n, tsgpu_g1* out) {
    return msm_device_batch(ctx, 1, &basis, &scalars, &n, out);
}

// window tables of a base array (device), allocated here; *table = nullptr when the mode is switched off
int build_tables(tsgpu_ctx* ctx, const g1_affine* pts, size_t n, g1_affine** table, unsigned* table_c) {
    *table = nullptr; *table_c = 0;
    if (!ctx->msm_tables || n == 0) return TSGPU_OK;
    const unsigned c = msm_table_window_bits(n), W = (255 + c - 1) / c;
    if ((size_t)W * n >= ((size_t)1 << 31)) return TSGPU_OK;   // entry indices are 31 bits: stay on per-window buckets
    g1_affine* t = nullptr;
    cudaError_t e = cudaMalloc((void**)&t, (size_t)W * n * sizeof(g1_affine));
    if (e != cudaSuccess) { cudaGetLastError(); return TSGPU_OK; }   // not enough memory for the tables: per-window buckets still work
    TempBuf cur;
    unsigned launches = 0;
    e = cur.alloc(n * sizeof(g1_xyzz), ctx->stream);
    if (e == cudaSuccess) e = msm_build_table(pts, n, c, t, cur.as<g1_xyzz>(), ctx->sm_count, ctx->stream, &launches);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    ctx->launches += launches;
    if (e != cudaSuccess) { cudaFree(t); return cuda_fail(ctx, e, "msm window tables"); }
    *table = t; *table_c = c;
    return TSGPU_OK;
}
MsmBasis srs_basis(const tsgpu_srs* srs) { return MsmBasis{srs->d, srs->n, srs->table, srs->table_c}; }

// byte-window table of the generator for k_fixed_base_mul: table[w * 255 + d - 1] = d * 256^w * G (affine); built once per process
const std::vector<g1_affine>& generator_table() {
    static std::vector<g1_affine> table;
    if (!table.empty()) return table;
    table.resize(32 * 255);
    G1J base = G1J::generator();
    std::vector<G1J> jac(32 * 255);
    for (int w = 0; w < 32; ++w) {
        G1J acc = base;
        for (int d = 1; d <= 255; ++d) { jac[w * 255 + d - 1] = acc; acc = acc.add(base); }
        base = acc;
    }
    // batch normalisation (Montgomery trick)
    std::vector<Fq64> pref(jac.size());
    Fq64 acc = Fq64::one();
    for (size_t i = 0; i < jac.size(); ++i) { pref[i] = acc; acc = acc * jac[i].z; }
    Fq64 inv = acc.inverse();
    for (size_t i = jac.size(); i-- > 0;) {
        Fq64 zi = inv * pref[i]; inv = inv * jac[i].z;
        Fq64 zi2 = zi.sqr();
        Fq64 ax = jac[i].x * zi2, ay = jac[i].y * zi2 * zi;
        memcpy(table[i].x.l, ax.l, 32); memcpy(table[i].y.l, ay.l, 32);
    }
    return table;
}
// out[i] = scalars[i] * G as affine points (scalars on the device, Montgomery form)
int fixed_base_points(tsgpu_ctx* ctx, const fr_t* scalars, size_t n, g1_affine* out) {
    if (!n) return TSGPU_OK;
    const std::vector<g1_affine>& table = generator_table();
    TempBuf dtable, xyzz;
    TSG_CUDA(ctx, dtable.alloc(table.size() * sizeof(g1_affine), ctx->stream));
    TSG_CUDA(ctx, xyzz.alloc(n * sizeof(g1_xyzz), ctx->stream));
    TSG_CUDA(ctx, cudaMemcpyAsync(dtable.p, table.data(), table.size() * sizeof(g1_affine), cudaMemcpyHostToDevice, ctx->stream));
    TSG_CUDA(ctx, launch_fixed_base_mul(scalars, n, dtable.as<g1_affine>(), xyzz.as<g1_xyzz>(), ctx->sm_count, ctx->stream));
    TSG_CUDA(ctx, launch_batch_to_affine(xyzz.as<g1_xyzz>(), n, out, ctx->sm_count, ctx->stream));
    ctx->launches += 2;
    return TSGPU_OK;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------- SRS
int tsgpu_srs_generate(tsgpu_ctx* ctx, const tsgpu_fr* tau, size_t n, tsgpu_srs** out) {
    int rc = tsgpu_srs_generate_range(ctx, tau, 0, n, out);
    if (!rc) { (*out)->has_tau = true; (*out)->tau = *tau; }
    return rc;
}

int tsgpu_srs_upload(tsgpu_ctx* ctx, const tsgpu_g1* powers, size_t n, tsgpu_srs** out) {
    if (!ctx || !out || (!powers && n)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    tsgpu_srs* srs = new (std::nothrow) tsgpu_srs;
    if (!srs) return fail(ctx, TSGPU_E_PROOF_GENERATION, "out of host memory");
    srs->n = n;
    cudaError_t e = cudaMalloc((void**)&srs->d, (n ? n : 1) * sizeof(g1_affine));
    if (e != cudaSuccess) { delete srs; return cuda_fail(ctx, e, "cudaMalloc(srs)"); }
    TempBuf jac, xyzz;
    TSG_CUDA(ctx, jac.alloc(n * sizeof(g1_jac), ctx->stream));
    TSG_CUDA(ctx, xyzz.alloc(n * sizeof(g1_xyzz), ctx->stream));
    if (n) {
        TSG_CUDA(ctx, cudaMemcpyAsync(jac.p, powers, n * sizeof(g1_jac), cudaMemcpyHostToDevice, ctx->stream));
        TSG_CUDA(ctx, launch_jac_to_xyzz(jac.as<g1_jac>(), n, xyzz.as<g1_xyzz>(), ctx->sm_count, ctx->stream));
        TSG_CUDA(ctx, launch_batch_to_affine(xyzz.as<g1_xyzz>(), n, srs->d, ctx->sm_count, ctx->stream));
        ctx->launches += 2;
    }
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int trc = build_tables(ctx, srs->d, n, &srs->table, &srs->table_c);
    if (trc) { tsgpu_srs_free(ctx, srs); return trc; }
    *out = srs;
    return TSGPU_OK;
}

int tsgpu_srs_download(tsgpu_ctx* ctx, const tsgpu_srs* srs, size_t first, size_t count, tsgpu_g1* out) {
    if (!ctx || !srs || (!out && count)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (first + count > srs->n) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "range exceeds SRS length");
    if (!count) return TSGPU_OK;
    TempBuf jac;
    TSG_CUDA(ctx, jac.alloc(count * sizeof(g1_jac), ctx->stream));
    TSG_CUDA(ctx, launch_affine_to_jac(srs->d + first, count, jac.as<g1_jac>(), ctx->sm_count, ctx->stream));
    ctx->launches += 1;
    TSG_CUDA(ctx, cudaMemcpyAsync(out, jac.p, count * sizeof(g1_jac), cudaMemcpyDeviceToHost, ctx->stream));
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TSGPU_OK;
}

size_t tsgpu_srs_len(const tsgpu_srs* srs) { return srs ? srs->n : 0; }
void tsgpu_srs_free(tsgpu_ctx* ctx, tsgpu_srs* srs) {
    (void)ctx;
    if (!srs) return;
    if (srs->d) cudaFree(srs->d);
    for (auto& kv : srs->lagrange) cudaFree(kv.second);
    for (auto& kv : srs->lagrange_table) cudaFree(kv.second.first);
    for (auto& kv : srs->lagrange_short_table) cudaFree(kv.second);
    for (auto& kv : srs->lagrange_slices) { cudaFree(kv.second.pts); if (kv.second.table) cudaFree(kv.second.table); if (kv.second.short_table) cudaFree(kv.second.short_table); }
    if (srs->table) cudaFree(srs->table);
    delete srs;
}

// ------------------------------------------------------------------------------------------- polynomials in HBM
int tsgpu_poly_upload(tsgpu_ctx* ctx, const tsgpu_fr* coeffs, size_t n, tsgpu_poly** out) {
    if (!ctx || !out || (!coeffs && n)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    tsgpu_poly* p = new (std::nothrow) tsgpu_poly;
    if (!p) return fail(ctx, TSGPU_E_PROOF_GENERATION, "out of host memory");
    p->n = n;
    cudaError_t e = cudaMallocAsync((void**)&p->d, (n ? n : 1) * sizeof(fr_t), ctx->stream);
    if (e != cudaSuccess) { delete p; return cuda_fail(ctx, e, "cudaMallocAsync(poly)"); }
    if (n) TSG_CUDA(ctx, cudaMemcpyAsync(p->d, coeffs, n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = p;
    return TSGPU_OK;
}
int tsgpu_poly_download(tsgpu_ctx* ctx, const tsgpu_poly* p, tsgpu_fr* out) {
    if (!ctx || !p || (!out && p->n)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (p->n) TSG_CUDA(ctx, cudaMemcpyAsync(out, p->d, p->n * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TSGPU_OK;
}
size_t tsgpu_poly_len(const tsgpu_poly* p) { return p ? p->n : 0; }
int tsgpu_poly_clone(tsgpu_ctx* ctx, const tsgpu_poly* p, tsgpu_poly** out) {
    if (!ctx || !p || !out) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    tsgpu_poly* c = new (std::nothrow) tsgpu_poly;
    if (!c) return fail(ctx, TSGPU_E_PROOF_GENERATION, "out of host memory");
    c->n = p->n;
    cudaError_t e = cudaMallocAsync((void**)&c->d, (p->n ? p->n : 1) * sizeof(fr_t), ctx->stream);
    if (e != cudaSuccess) { delete c; return cuda_fail(ctx, e, "cudaMallocAsync(poly)"); }
    if (p->n) TSG_CUDA(ctx, cudaMemcpyAsync(c->d, p->d, p->n * sizeof(fr_t), cudaMemcpyDeviceToDevice, ctx->stream));
    c->short64 = p->short64;
    *out = c;
    return TSGPU_OK;
}
void tsgpu_poly_free(tsgpu_ctx* ctx, tsgpu_poly* p) {
    if (!p) return;
    if (p->ready) { if (ctx) cudaStreamWaitEvent(ctx->stream, p->ready, 0); else cudaEventSynchronize(p->ready); cudaEventDestroy(p->ready); }
    if (p->d) cudaFreeAsync(p->d, ctx ? ctx->stream : nullptr);
    delete p;
}

// values[i] = Fr::from(v[i]) for i < n, zero-padded to `padded` entries (twist.rs:115-122,141-148)
int tsgpu_poly_from_u64(tsgpu_ctx* ctx, const uint64_t* v, size_t n, size_t padded, tsgpu_poly** out) {
    if (!ctx || !out || (!v && n) || padded < n) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "bad argument");
    tsgpu_poly* p = new (std::nothrow) tsgpu_poly;
    if (!p) return fail(ctx, TSGPU_E_PROOF_GENERATION, "out of host memory");
    p->n = padded;
    cudaError_t e = cudaMallocAsync((void**)&p->d, (padded ? padded : 1) * sizeof(fr_t), ctx->stream);
    if (e != cudaSuccess) { delete p; return cuda_fail(ctx, e, "cudaMallocAsync(poly)"); }
    if (padded > n) TSG_CUDA(ctx, cudaMemsetAsync(p->d + n, 0, (padded - n) * sizeof(fr_t), ctx->stream));
    TempBuf src;
    TSG_CUDA(ctx, src.alloc(n * 8, ctx->stream));
    if (n) {
        TSG_CUDA(ctx, cudaMemcpyAsync(src.p, v, n * 8, cudaMemcpyHostToDevice, ctx->stream));
        TSG_CUDA(ctx, launch_fr_from_u64(src.as<unsigned long long>(), n, p->d, 0, 0, ctx->sm_count, ctx->stream));
        ctx->launches += 1;
    }
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    p->short64 = true;
    *out = p;
    return TSGPU_OK;
}
// host values, zero-padded to `padded` entries (Vec::resize, twist.rs:146-148)
int tsgpu_poly_upload_padded(tsgpu_ctx* ctx, const tsgpu_fr* vals, size_t n, size_t padded, tsgpu_poly** out) {
    if (!ctx || !out || (!vals && n) || padded < n) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "bad argument");
    tsgpu_poly* p = new (std::nothrow) tsgpu_poly;
    if (!p) return fail(ctx, TSGPU_E_PROOF_GENERATION, "out of host memory");
    p->n = padded;
    cudaError_t e = cudaMallocAsync((void**)&p->d, (padded ? padded : 1) * sizeof(fr_t), ctx->stream);
    if (e != cudaSuccess) { delete p; return cuda_fail(ctx, e, "cudaMallocAsync(poly)"); }
    if (padded > n) TSG_CUDA(ctx, cudaMemsetAsync(p->d + n, 0, (padded - n) * sizeof(fr_t), ctx->stream));
    if (n) TSG_CUDA(ctx, cudaMemcpyAsync(p->d, vals, n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = p;
    return TSGPU_OK;
}

// same vector, but the host -> device copy runs on the context's side stream and the call returns without waiting for it: whatever is enqueued
// on the context afterwards overlaps the transfer.  Nothing in the library waits for the copy by itself - call tsgpu_poly_wait before the first use.
int tsgpu_poly_upload_padded_async(tsgpu_ctx* ctx, const tsgpu_fr* vals, size_t n, size_t padded, tsgpu_poly** out) {
    if (!ctx || !out || (!vals && n) || padded < n) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "bad argument");
    if (!ctx->copy_stream) TSG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    tsgpu_poly* p = new (std::nothrow) tsgpu_poly;
    if (!p) return fail(ctx, TSGPU_E_PROOF_GENERATION, "out of host memory");
    p->n = padded;
    cudaError_t e = cudaMallocAsync((void**)&p->d, (padded ? padded : 1) * sizeof(fr_t), ctx->stream);
    if (e != cudaSuccess) { delete p; return cuda_fail(ctx, e, "cudaMallocAsync(poly)"); }
    cudaEvent_t allocated = nullptr;
    e = cudaEventCreateWithFlags(&allocated, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(allocated, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, allocated, 0);     // the side stream may touch the block once the allocation is ordered
    if (e == cudaSuccess && padded > n) e = cudaMemsetAsync(p->d + n, 0, (padded - n) * sizeof(fr_t), ctx->copy_stream);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(p->d, vals, n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->copy_stream);
    if (e == cudaSuccess) e = cudaEventRecord(p->ready, ctx->copy_stream);
    if (allocated) cudaEventDestroy(allocated);
    if (e != cudaSuccess) { tsgpu_poly_free(ctx, p); return cuda_fail(ctx, e, "side-stream upload"); }
    *out = p;
    return TSGPU_OK;
}
int tsgpu_poly_wait(tsgpu_ctx* ctx, tsgpu_poly* p) {
    if (!ctx || !p) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (!p->ready) return TSGPU_OK;
    TSG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, p->ready, 0));
    cudaEventDestroy(p->ready); p->ready = nullptr;
    return TSGPU_OK;
}
int tsgpu_poly_in_flight(const tsgpu_poly* p) { return p && p->ready ? 1 : 0; }

// ------------------------------------------------------------------------------------------- interpolation
static int log2_exact(size_t n) { int l = 0; while (((size_t)1 << l) < n) ++l; return ((size_t)1 << l) == n ? l : -1; }

int tsgpu_interpolate_prepare(tsgpu_ctx* ctx, unsigned log_n) {
    if (!ctx) return TSGPU_E_INVALID_PARAMETERS;
    if (log_n > 27) return fail(ctx, TSGPU_E_POLYNOMIAL, "interpolation size exceeds the 2^28 two-adicity of Fr");
    TSG_CUDA(ctx, interp_prepare(ctx, log_n));
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TSGPU_OK;
}
// in place: values at x = 0..n-1  ->  monomial coefficients (poly_utils::lagrange_interpolate on x_i = i)
int tsgpu_poly_interpolate_iota(tsgpu_ctx* ctx, tsgpu_poly* p) {
    if (!ctx || !p) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (p->n == 0) return TSGPU_OK;   // lagrange_interpolate(&[]) = [] (polynomials.rs:303-305)
    int l = log2_exact(p->n);
    if (l < 0) return fail(ctx, TSGPU_E_POLYNOMIAL, "interpolation length must be a power of two (Twist/Shout pad first)");
    if (l > 27) return fail(ctx, TSGPU_E_POLYNOMIAL, "interpolation size exceeds the 2^28 two-adicity of Fr");
    KernelTimer kt(ctx, "interpolate");
    p->short64 = false;   // coefficients are field-sized
    TSG_CUDA(ctx, interp_run(ctx, p->d, (unsigned)l, p->d));
    return TSGPU_OK;
}
// lagrange_interpolate on (i, values[i]), i < n, for ANY n: the Newton coefficients of the n points, zero-extended to the next power of
// two, are converted to the monomial basis at that size (higher Newton terms being zero does not change the polynomial); n coefficients back
static int interpolate_any_dev(tsgpu_ctx* ctx, const tsgpu_fr* values, size_t n, tsgpu_poly** out) {
    size_t N = 1; while (N < n) N <<= 1;
    int l = log2_exact(N);
    if (l > 27) return fail(ctx, TSGPU_E_POLYNOMIAL, "interpolation size exceeds the 2^28 two-adicity of Fr");
    tsgpu_poly* p = nullptr;
    int rc = tsgpu_poly_upload_padded(ctx, values, n, N, &p);
    if (rc) return rc;
    {
        KernelTimer kt(ctx, "interpolate");
        cudaError_t e = interp_run(ctx, p->d, (unsigned)l, p->d, n);
        if (e != cudaSuccess) { tsgpu_poly_free(ctx, p); return cuda_fail(ctx, e, "interpolation"); }
    }
    p->n = n;       // degree < n: the tail of the buffer is zero
    *out = p;
    return TSGPU_OK;
}
int tsgpu_interpolate_iota(tsgpu_ctx* ctx, const tsgpu_fr* values, size_t n, tsgpu_fr* coeffs) {
    if (!ctx || (n && (!values || !coeffs))) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (n == 0) return TSGPU_OK;   // lagrange_interpolate(&[]) = [] (polynomials.rs:303-305)
    tsgpu_poly* p = nullptr;
    int rc = interpolate_any_dev(ctx, values, n, &p);
    if (rc) return rc;
    rc = tsgpu_poly_download(ctx, p, coeffs);
    tsgpu_poly_free(ctx, p);
    return rc;
}

// ------------------------------------------------------------------------------------------- MSM / KZG
int tsgpu_msm_g1(tsgpu_ctx* ctx, const tsgpu_g1a* bases, const tsgpu_fr* scalars, size_t n, tsgpu_g1* out) {
    if (!ctx || !out || ((!bases || !scalars) && n)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    TempBuf b, s;
    TSG_CUDA(ctx, b.alloc(n * sizeof(g1_affine), ctx->stream));
    TSG_CUDA(ctx, s.alloc(n * sizeof(fr_t), ctx->stream));
    if (n) {
        TSG_CUDA(ctx, cudaMemcpyAsync(b.p, bases, n * sizeof(g1_affine), cudaMemcpyHostToDevice, ctx->stream));
        TSG_CUDA(ctx, cudaMemcpyAsync(s.p, scalars, n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    return msm_device(ctx, MsmBasis{b.as<g1_affine>(), n, nullptr, 0}, s.as<fr_t>(), n, out);
}

// `count` commitments over the same SRS in one MSM pass (the two commitments of a Twist / Shout proof)
int tsgpu_kzg_commit_batch_dev(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_poly* const* polys, size_t count, tsgpu_g1* outs) {
    if (!ctx || !srs || (count && (!polys || !outs))) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    for (size_t i = 0; i < count; ++i) {
        if (!polys[i]) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
        if (polys[i]->n > srs->n) return fail(ctx, TSGPU_E_COMMITMENT, "Polynomial degree exceeds setup size");   // commitments.rs:166-170
    }
    for (size_t i0 = 0; i0 < count; i0 += MSM_MAX_BATCH) {
        const int K = (int)(count - i0 < (size_t)MSM_MAX_BATCH ? count - i0 : MSM_MAX_BATCH);
        MsmBasis basis[MSM_MAX_BATCH]; const fr_t* sc[MSM_MAX_BATCH]; size_t n[MSM_MAX_BATCH];
        for (int k = 0; k < K; ++k) { basis[k] = srs_basis(srs); sc[k] = polys[i0 + k]->d; n[k] = polys[i0 + k]->n; }
        int rc = msm_device_batch(ctx, K, basis, sc, n, outs + i0);
        if (rc) return rc;
    }
    return TSGPU_OK;
}
int tsgpu_kzg_commit_dev(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_poly* poly, tsgpu_g1* out) {
    if (!poly || !out) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    return tsgpu_kzg_commit_batch_dev(ctx, srs, &poly, 1, out);
}

int tsgpu_kzg_commit(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_fr* polynomial, size_t n, tsgpu_g1* out) {
    if (!ctx || !srs || !out || (!polynomial && n)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (n > srs->n) return fail(ctx, TSGPU_E_COMMITMENT, "Polynomial degree exceeds setup size");
    TempBuf s;
    TSG_CUDA(ctx, s.alloc(n * sizeof(fr_t), ctx->stream));
    if (n) TSG_CUDA(ctx, cudaMemcpyAsync(s.p, polynomial, n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    return msm_device(ctx, srs_basis(srs), s.as<fr_t>(), n, out);
}

// value = P(z), proof = commit((P - value) / (x - z))   (commitments.rs:182-199)
// `count` (<= MSM_MAX_BATCH) openings at the same point: the value / quotient scans run back to back, the quotient commitments as one MSM pass
int tsgpu_kzg_open_batch_dev(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_poly* const* polys, size_t count, const tsgpu_fr* z, tsgpu_fr* values, tsgpu_g1* proofs) {
    if (!ctx || !srs || !z || (count && (!polys || !values || !proofs))) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (count > (size_t)MSM_MAX_BATCH) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "too many openings in one batch");
    size_t total = 0;
    for (size_t i = 0; i < count; ++i) {
        if (!polys[i]) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
        if (polys[i]->n && polys[i]->n - 1 > srs->n) return fail(ctx, TSGPU_E_COMMITMENT, "Polynomial degree exceeds setup size");
        total += polys[i]->n;
    }
    // scan weights on the host: pw[s] = z^(SPAN 2^s), W = z^(THREADS SPAN)
    Fr64 zz = Fr64::from_raw(z->l);
    Fr64 w = zz;
    for (size_t k = 1; k < POLY_SPAN; k <<= 1) w = w.sqr();   // z^SPAN (SPAN is a power of two)
    fr_t pw[POLY_PW + 1];
    for (int s = 0; s <= POLY_PW; ++s) { memcpy(pw[s].l, w.l, 32); w = w.sqr(); }
    TempBuf dpw, totals, carry, val;
    TSG_CUDA(ctx, dpw.alloc(sizeof(pw), ctx->stream));
    TSG_CUDA(ctx, val.alloc(MSM_MAX_BATCH * sizeof(fr_t), ctx->stream));
    cudaError_t aerr;
    fr_t* q_all = (fr_t*)arena_get(ctx, tsgpu_ctx::ARENA_QUOT, (total + 1) * sizeof(fr_t), &aerr);
    if (!q_all) return cuda_fail(ctx, aerr, "cudaMalloc(quotient)");
    TSG_CUDA(ctx, cudaMemcpyAsync(dpw.p, pw, sizeof(pw), cudaMemcpyHostToDevice, ctx->stream));
    TSG_CUDA(ctx, cudaMemsetAsync(val.p, 0, MSM_MAX_BATCH * sizeof(fr_t), ctx->stream));   // empty polynomial: value 0, empty quotient (commitments.rs:306-308, 322-324)
    fr_t zf; memcpy(zf.l, z->l, 32);
    MsmBasis basis[MSM_MAX_BATCH]; const fr_t* sc[MSM_MAX_BATCH]; size_t qn[MSM_MAX_BATCH];
    size_t off = 0;
    for (size_t i = 0; i < count; ++i) {
        const size_t n = polys[i]->n;
        basis[i] = srs_basis(srs); sc[i] = q_all + off; qn[i] = n ? n - 1 : 0;
        if (n) {
            const size_t nblocks = poly_num_blocks(n);
            TSG_CUDA(ctx, totals.alloc(nblocks * sizeof(fr_t), ctx->stream));
            TSG_CUDA(ctx, carry.alloc(nblocks * sizeof(fr_t), ctx->stream));
            unsigned launches = 0;
            KernelTimer kt_open(ctx, "open_scan");
            TSG_CUDA(ctx, poly_open_launch(polys[i]->d, n, zf, dpw.as<fr_t>(), pw[POLY_PW], totals.as<fr_t>(), carry.as<fr_t>(), q_all + off, val.as<fr_t>() + i,
                                           ctx->stream, &launches));
            ctx->launches += launches;
            cudaFreeAsync(totals.p, ctx->stream); totals.p = nullptr;
            cudaFreeAsync(carry.p, ctx->stream); carry.p = nullptr;
        }
        off += n;
    }
    TSG_CUDA(ctx, cudaMemcpyAsync(ctx->host_out, val.p, count * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    int rc = msm_device_batch(ctx, (int)count, basis, sc, qn, proofs);
    if (rc) return rc;
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(values, ctx->host_out, count * 32);
    return TSGPU_OK;
}
int tsgpu_kzg_open_dev(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_poly* poly, const tsgpu_fr* z, tsgpu_fr* value, tsgpu_g1* proof) {
    if (!poly || !value || !proof) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    return tsgpu_kzg_open_batch_dev(ctx, srs, &poly, 1, z, value, proof);
}

int tsgpu_kzg_open(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_fr* polynomial, size_t n, const tsgpu_fr* z, tsgpu_fr* value, tsgpu_g1* proof) {
    tsgpu_poly* p = nullptr;
    int rc = tsgpu_poly_upload(ctx, polynomial, n, &p);
    if (rc) return rc;
    rc = tsgpu_kzg_open_dev(ctx, srs, p, z, value, proof);
    tsgpu_poly_free(ctx, p);
    return rc;
}

// ---- host-side point helpers (CPU only): what the transcript and proof bytes need
void tsgpu_g1_hash(const tsgpu_g1* p, tsgpu_fr* out) {            // KZGCommitmentValue::hash, commitments.rs:73-84
    G1J j; memcpy(&j, p, 96);
    Fr64 h = tsg::host::g1_hash(j);
    memcpy(out->l, h.l, 32);
}
void tsgpu_g1_compress(const tsgpu_g1* p, uint8_t out[32]) {      // ark-serialize compressed (commitments.rs:106-118)
    G1J j; memcpy(&j, p, 96);
    tsg::host::g1_compress(j, out);
}
int tsgpu_g1_equal(const tsgpu_g1* a, const tsgpu_g1* b) {
    G1J x, y; memcpy(&x, a, 96); memcpy(&y, b, 96);
    return x.equals(y) ? 1 : 0;
}

}  // extern "C"

extern "C" {
// g1_powers[first .. first + n) of setup_params (utils.rs:89-96): the slice a point-sharded MSM rank needs
int tsgpu_srs_generate_range(tsgpu_ctx* ctx, const tsgpu_fr* tau, size_t first, size_t n, tsgpu_srs** out) {
    if (!ctx || !tau || !out) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    tsgpu_srs* srs = new (std::nothrow) tsgpu_srs;
    if (!srs) return fail(ctx, TSGPU_E_PROOF_GENERATION, "out of host memory");
    srs->n = n;
    cudaError_t e = cudaMalloc((void**)&srs->d, (n ? n : 1) * sizeof(g1_affine));
    if (e != cudaSuccess) { delete srs; return cuda_fail(ctx, e, "cudaMalloc(srs)"); }
    TempBuf scal;
    fr_t t; memcpy(t.l, tau->l, 32);
    int rc = TSGPU_OK;
    if (n) {
        cudaError_t ce = scal.alloc(n * sizeof(fr_t), ctx->stream);
        if (ce == cudaSuccess) ce = launch_tau_powers(t, first, n, scal.as<fr_t>(), ctx->sm_count, ctx->stream);   // tau^(first + i)
        if (ce != cudaSuccess) rc = cuda_fail(ctx, ce, "tau powers");
        ctx->launches += 1;
        if (!rc) rc = fixed_base_points(ctx, scal.as<fr_t>(), n, srs->d);
    }
    if (!rc) { cudaError_t ce = cudaStreamSynchronize(ctx->stream); if (ce != cudaSuccess) rc = cuda_fail(ctx, ce, "srs generation"); }
    if (!rc) rc = build_tables(ctx, srs->d, n, &srs->table, &srs->table_c);
    if (rc) { tsgpu_srs_free(ctx, srs); return rc; }
    *out = srs;
    return TSGPU_OK;
}

// ------------------------------------------------------------------------------------------- evaluation-basis KZG
static g1_affine* lagrange_basis_unlocked(const tsgpu_srs* srs, size_t m) {
    auto it = srs->lagrange.find(m);
    return it == srs->lagrange.end() ? nullptr : it->second;
}
static g1_affine* lagrange_basis(const tsgpu_srs* srs, size_t m) {
    std::lock_guard<std::mutex> lock(srs->cache_mu);
    return lagrange_basis_unlocked(srs, m);
}
static MsmBasis lagrange_msm_basis(const tsgpu_srs* srs, size_t m) {
    std::lock_guard<std::mutex> lock(srs->cache_mu);
    auto it = srs->lagrange_table.find(m);
    MsmBasis b{lagrange_basis_unlocked(srs, m), m, it == srs->lagrange_table.end() ? nullptr : it->second.first, it == srs->lagrange_table.end() ? 0u : it->second.second};
    // short scalars (< 2^64): a dedicated table of MSM_SHORT_C-bit windows when the full table's windows are wider, else the low
    // windows of the full table itself
    auto sh = srs->lagrange_short_table.find(m);
    if (sh != srs->lagrange_short_table.end()) { b.short_table = sh->second; b.short_c = MSM_SHORT_C; b.short_windows = MSM_SHORT_WINDOWS; }
    else if (b.table && b.table_c <= MSM_SHORT_C) { b.short_table = b.table; b.short_c = b.table_c; b.short_windows = (65 + b.table_c - 1) / b.table_c; }
    return b;
}
int tsgpu_srs_has_lagrange(const tsgpu_srs* srs, size_t m) { return srs && lagrange_basis(srs, m) ? 1 : 0; }
int tsgpu_srs_can_lagrange(const tsgpu_srs* srs) { return srs && srs->has_tau ? 1 : 0; }

// [L_j(tau)]_1 for the nodes first .. first + count - 1 of the m-node domain (m a power of two, m <= SRS length):
// L_j(tau) = N(tau) w_j / (tau - j) on the device, then the same fixed-base kernel that builds g1_powers, then the window tables.
static int build_lagrange_range(tsgpu_ctx* ctx, tsgpu_srs* srs, size_t m, size_t first, size_t count, tsgpu_srs::Slice* out) {
    if (!srs->has_tau) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "this SRS was uploaded without its trapdoor: no evaluation basis");
    int lg = log2_exact(m);
    if (lg < 0 || m > srs->n || first + count > m || count == 0) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "evaluation-basis size must be a power of two within the SRS");
    Fr64 tcan = Fr64::from_raw(srs->tau.l).from_mont();
    if (!tcan.l[1] && !tcan.l[2] && !tcan.l[3] && tcan.l[0] < m) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "tau is an interpolation node");
    const fr_t* ifact = nullptr;
    TSG_CUDA(ctx, interp_factorials(ctx, (unsigned)lg, &ifact));
    g1_affine* basis = nullptr;
    cudaError_t e = cudaMalloc((void**)&basis, count * sizeof(g1_affine));
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(lagrange basis)");
    TempBuf inv, scratch, scal, prod;
    int rc = TSGPU_OK;
    fr_t t; memcpy(t.l, srs->tau.l, 32);
    fr_t ntau;
    unsigned launches = 0;
    // inverses over the WHOLE domain (N(tau) is a product over all nodes; 0.2 ms at 2^20), scalars and points only for the range
    cudaError_t ce = inv.alloc(m * sizeof(fr_t), ctx->stream);
    if (ce == cudaSuccess) ce = scratch.alloc(lag_binv_scratch(m) * sizeof(fr_t), ctx->stream);
    if (ce == cudaSuccess) ce = scal.alloc(count * sizeof(fr_t), ctx->stream);
    if (ce == cudaSuccess) ce = prod.alloc(sizeof(fr_t), ctx->stream);
    if (ce == cudaSuccess) ce = launch_node_inverses(t, m, inv.as<fr_t>(), scratch.as<fr_t>(), ctx->host_scratch, &ntau, ctx->sm_count, ctx->stream, &launches);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(prod.p, &ntau, sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream);
    if (ce == cudaSuccess) ce = launch_lagrange_scalars(inv.as<fr_t>() + first, ifact, prod.as<fr_t>(), m, first, count, scal.as<fr_t>(), ctx->sm_count, ctx->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);   // ntau is a stack variable
    ctx->launches += launches + 1;
    if (ce != cudaSuccess) rc = cuda_fail(ctx, ce, "lagrange scalars");
    if (!rc) rc = fixed_base_points(ctx, scal.as<fr_t>(), count, basis);
    if (!rc) { ce = cudaStreamSynchronize(ctx->stream); if (ce != cudaSuccess) rc = cuda_fail(ctx, ce, "lagrange basis"); }
    if (rc) { cudaFree(basis); return rc; }
    out->pts = basis;
    if ((rc = build_tables(ctx, basis, count, &out->table, &out->table_c))) { cudaFree(basis); out->pts = nullptr; return rc; }
    if (out->table && out->table_c > MSM_SHORT_C) {   // dedicated short-scalar table when the full table's bucket set is the larger one
        g1_affine* st = nullptr;
        if (cudaMalloc((void**)&st, (size_t)MSM_SHORT_WINDOWS * count * sizeof(g1_affine)) == cudaSuccess) {
            TempBuf cur; unsigned l2 = 0;
            cudaError_t e2 = cur.alloc(count * sizeof(g1_xyzz), ctx->stream);
            if (e2 == cudaSuccess) e2 = msm_build_table(basis, count, MSM_SHORT_C, st, cur.as<g1_xyzz>(), ctx->sm_count, ctx->stream, &l2, MSM_SHORT_WINDOWS);
            if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(ctx->stream);
            ctx->launches += l2;
            if (e2 != cudaSuccess) { cudaFree(st); return cuda_fail(ctx, e2, "short window tables"); }
            out->short_table = st;
        } else cudaGetLastError();
    }
    return TSGPU_OK;
}
// the whole basis of an m-node domain; the handle caches the result (a const handle is a cache here)
int tsgpu_srs_lagrange_prepare(tsgpu_ctx* ctx, const tsgpu_srs* srs_c, size_t m) {
    if (!ctx || !srs_c) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    tsgpu_srs* srs = const_cast<tsgpu_srs*>(srs_c);
    std::lock_guard<std::mutex> lock(srs->cache_mu);
    if (lagrange_basis_unlocked(srs, m)) return TSGPU_OK;
    tsgpu_srs::Slice sl;
    int rc = build_lagrange_range(ctx, srs, m, 0, m, &sl);
    if (rc) return rc;
    srs->lagrange[m] = sl.pts;
    if (sl.table) srs->lagrange_table[m] = {sl.table, sl.table_c};
    if (sl.short_table) srs->lagrange_short_table[m] = sl.short_table;
    return TSGPU_OK;
}
// one rank's slice of it (sharded proving): nodes first .. first + count - 1
int tsgpu_srs_lagrange_prepare_range(tsgpu_ctx* ctx, const tsgpu_srs* srs_c, size_t m, size_t first, size_t count) {
    if (!ctx || !srs_c) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    tsgpu_srs* srs = const_cast<tsgpu_srs*>(srs_c);
    std::lock_guard<std::mutex> lock(srs->cache_mu);
    auto key = std::make_tuple(m, first, count);
    if (srs->lagrange_slices.count(key)) return TSGPU_OK;
    tsgpu_srs::Slice sl;
    int rc = build_lagrange_range(ctx, srs, m, first, count, &sl);
    if (rc) return rc;
    srs->lagrange_slices[key] = sl;
    return TSGPU_OK;
}
static MsmBasis lagrange_slice_basis(const tsgpu_srs* srs, size_t m, size_t first, size_t count) {
    std::lock_guard<std::mutex> lock(srs->cache_mu);
    const tsgpu_srs::Slice& sl = srs->lagrange_slices.at(std::make_tuple(m, first, count));
    MsmBasis b{sl.pts, count, sl.table, sl.table_c};
    if (sl.short_table) { b.short_table = sl.short_table; b.short_c = MSM_SHORT_C; b.short_windows = MSM_SHORT_WINDOWS; }
    else if (b.table && b.table_c <= MSM_SHORT_C) { b.short_table = b.table; b.short_c = b.table_c; b.short_windows = (65 + b.table_c - 1) / b.table_c; }
    return b;
}

// ---- sharded evaluation-basis commit / open: each rank holds the nodes [first, first + count) of the m-node domain -----------------
// partial commitments of `k` value slices over this rank's basis slice (one batched MSM pass); the caller sums them over the ranks
int tsgpu_kzg_commit_values_slice_batch_dev(tsgpu_ctx* ctx, const tsgpu_srs* srs, size_t m, size_t first, const tsgpu_poly* const* slices, size_t k, tsgpu_g1* partials) {
    if (!ctx || !srs || !slices || !partials || k == 0 || k > (size_t)MSM_MAX_BATCH) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "bad argument");
    const size_t count = slices[0]->n;
    for (size_t i = 0; i < k; ++i) if (!slices[i] || slices[i]->n != count) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "slices must have equal length");
    int rc = tsgpu_srs_lagrange_prepare_range(ctx, srs, m, first, count);
    if (rc) return rc;
    MsmBasis basis[MSM_MAX_BATCH]; const fr_t* sc[MSM_MAX_BATCH]; size_t n[MSM_MAX_BATCH];
    for (size_t i = 0; i < k; ++i) { basis[i] = lagrange_slice_basis(srs, m, first, count); sc[i] = slices[i]->d; n[i] = count; }
    return msm_device_batch(ctx, (int)k, basis, sc, n, partials, /*maybe_short=*/true);
}
// opening, phase 1: out[0] = prod over this rank's nodes of (z - j); out[1 + i] = sum over this rank's nodes of w_j v_j / (z - j) for slice i.
// The caller multiplies the products of all ranks (= N(z)), adds the partial sums of all ranks and gets value_i = N(z) * sum_i.
int tsgpu_kzg_open_values_slice_partial(tsgpu_ctx* ctx, size_t m, size_t first, const tsgpu_poly* const* slices, size_t k, const tsgpu_fr* z, tsgpu_fr* out) {
    if (!ctx || !slices || !z || !out || k == 0 || k > (size_t)MSM_MAX_BATCH) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "bad argument");
    const size_t count = slices[0]->n;
    int lg = log2_exact(m);
    if (lg < 0 || first + count > m) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "bad slice");
    Fr64 zcan = Fr64::from_raw(z->l).from_mont();
    if (!zcan.l[1] && !zcan.l[2] && !zcan.l[3] && zcan.l[0] < m) return fail(ctx, TSGPU_E_POLYNOMIAL, "opening point is an interpolation node");
    const fr_t* ifact = nullptr;
    TSG_CUDA(ctx, interp_factorials(ctx, (unsigned)lg, &ifact));
    cudaError_t aerr;
    fr_t* inv = (fr_t*)arena_get(ctx, tsgpu_ctx::ARENA_BARY, (count + lag_binv_scratch(count) + 8) * sizeof(fr_t), &aerr);
    if (!inv) return cuda_fail(ctx, aerr, "cudaMalloc(node inverses)");
    ctx->bary_inv = inv; ctx->bary_spans = inv + count; ctx->bary_n = count; ctx->bary_z = *z;
    // 1 / (z - first - j): the batch inversion counts nodes from 0, so it is run at the shifted point z - first
    Fr64 zs = Fr64::from_raw(z->l) - Fr64::from_u64((uint64_t)first);
    fr_t zf; memcpy(zf.l, zs.l, 32);
    fr_t slice_prod;
    unsigned launches = 0;
    KernelTimer kt(ctx, "open_bary");
    TSG_CUDA(ctx, launch_node_inverses(zf, count, inv, (fr_t*)ctx->bary_spans, ctx->host_scratch, &slice_prod, ctx->sm_count, ctx->stream, &launches));
    ctx->launches += launches;
    memcpy(out[0].l, slice_prod.l, 32);
    fr_t* one_dev = (fr_t*)ctx->bary_spans + lag_binv_scratch(count);
    fr_t one = fr_t::one();
    memcpy(ctx->host_scratch + 41, &one, sizeof(fr_t));
    TSG_CUDA(ctx, cudaMemcpyAsync(one_dev, ctx->host_scratch + 41, sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    for (size_t i = 0; i < k; ++i) {
        if (!slices[i] || slices[i]->n != count) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "slices must have equal length");
        TSG_CUDA(ctx, launch_bary_partial(slices[i]->d, inv, ifact, m, first, count, one_dev, ctx->partials, ctx->ticket, ctx->dev_out + i, ctx->sm_count, ctx->stream));
        ctx->launches += 1;
    }
    TSG_CUDA(ctx, cudaMemcpyAsync(ctx->host_out, ctx->dev_out, k * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(out + 1, ctx->host_out, k * sizeof(fr_t));
    return TSGPU_OK;
}
// opening, phase 2 (after phase 1 with the same z on this context): quotient values over this rank's nodes for the global values[i],
// committed over the basis slice in one batched pass; the caller sums the partial proofs over the ranks
int tsgpu_kzg_open_values_slice_finish(tsgpu_ctx* ctx, const tsgpu_srs* srs, size_t m, size_t first, const tsgpu_poly* const* slices, size_t k,
                                       const tsgpu_fr* values, tsgpu_g1* partial_proofs) {
    if (!ctx || !srs || !slices || !values || !partial_proofs || k == 0 || k > (size_t)MSM_MAX_BATCH) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "bad argument");
    const size_t count = slices[0]->n;
    if (!ctx->bary_inv || ctx->bary_n != count) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "open_values_slice_partial must run first");
    int rc = tsgpu_srs_lagrange_prepare_range(ctx, srs, m, first, count);
    if (rc) return rc;
    cudaError_t aerr;
    fr_t* q_all = (fr_t*)arena_get(ctx, tsgpu_ctx::ARENA_QUOT, (k * count + MSM_MAX_BATCH) * sizeof(fr_t), &aerr);
    if (!q_all) return cuda_fail(ctx, aerr, "cudaMalloc(quotient)");
    fr_t* val = q_all + k * count;
    memcpy(ctx->host_scratch + 44, values, k * sizeof(fr_t));
    TSG_CUDA(ctx, cudaMemcpyAsync(val, ctx->host_scratch + 44, k * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    MsmBasis basis[MSM_MAX_BATCH]; const fr_t* sc[MSM_MAX_BATCH]; size_t n[MSM_MAX_BATCH];
    for (size_t i = 0; i < k; ++i) {
        TSG_CUDA(ctx, launch_bary_quotient(slices[i]->d, (const fr_t*)ctx->bary_inv, val + i, count, q_all + i * count, ctx->sm_count, ctx->stream));
        ctx->launches += 1;
        basis[i] = lagrange_slice_basis(srs, m, first, count); sc[i] = q_all + i * count; n[i] = count;
    }
    return msm_device_batch(ctx, (int)k, basis, sc, n, partial_proofs);
}

// commit(interpolant of values on 0..m-1) = sum_j values[j] * [L_j(tau)]_1 : vector_to_polynomial + commit in one MSM.
// `count` vectors (their lengths may differ: each uses the basis of its own length) in one MSM pass.
int tsgpu_kzg_commit_values_batch_dev(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_poly* const* values, size_t count, tsgpu_g1* outs) {
    if (!ctx || !srs || (count && (!values || !outs))) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (count > (size_t)MSM_MAX_BATCH) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "too many commitments in one batch");
    MsmBasis basis[MSM_MAX_BATCH]; const fr_t* sc[MSM_MAX_BATCH]; size_t n[MSM_MAX_BATCH];
    for (size_t i = 0; i < count; ++i) {
        if (!values[i]) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
        if (values[i]->n > srs->n) return fail(ctx, TSGPU_E_COMMITMENT, "Polynomial degree exceeds setup size");
        n[i] = values[i]->n; sc[i] = values[i]->d;
        basis[i] = MsmBasis{nullptr, 0, nullptr, 0};
        if (n[i]) {
            int rc = tsgpu_srs_lagrange_prepare(ctx, srs, n[i]);
            if (rc) return rc;
            basis[i] = lagrange_msm_basis(srs, n[i]);
        }
    }
    bool known_short = count > 0;
    for (size_t i = 0; i < count; ++i) known_short = known_short && values[i]->short64;
    return msm_device_batch(ctx, (int)count, basis, sc, n, outs, /*maybe_short=*/true, known_short);
}
int tsgpu_kzg_commit_values_dev(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_poly* values, tsgpu_g1* out) {
    if (!values || !out) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    return tsgpu_kzg_commit_values_batch_dev(ctx, srs, &values, 1, out);
}

// KZGCommitment::open on the interpolants of `count` value vectors at one point: value = P(z) by the barycentric formula,
// proof = commitment to the quotient through its values Q(j) = (v_j - value) / (j - z), the quotient commitments as one MSM
// pass.  z must not be one of the nodes (TSGPU_E_POLYNOMIAL).  The node inverses 1/(z - j) are computed once for the longest vector.
int tsgpu_kzg_open_values_batch_dev(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_poly* const* values, size_t count, const tsgpu_fr* z,
                                    tsgpu_fr* out_values, tsgpu_g1* proofs) {
    if (!ctx || !srs || !z || (count && (!values || !out_values || !proofs))) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (count > (size_t)MSM_MAX_BATCH) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "too many openings in one batch");
    size_t mmax = 0, total = 0;
    for (size_t i = 0; i < count; ++i) {
        if (!values[i]) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
        mmax = values[i]->n > mmax ? values[i]->n : mmax; total += values[i]->n;
    }
    if (mmax == 0) {
        G1J id = G1J::identity();
        for (size_t i = 0; i < count; ++i) { memset(&out_values[i], 0, 32); memcpy(&proofs[i], &id, 96); }
        return TSGPU_OK;
    }
    Fr64 zcan = Fr64::from_raw(z->l).from_mont();
    if (!zcan.l[1] && !zcan.l[2] && !zcan.l[3] && zcan.l[0] < mmax) return fail(ctx, TSGPU_E_POLYNOMIAL, "opening point is an interpolation node");
    const fr_t* ifact = nullptr;
    TSG_CUDA(ctx, interp_factorials(ctx, (unsigned)log2_exact(mmax), &ifact));
    cudaError_t aerr;
    fr_t* q_all = (fr_t*)arena_get(ctx, tsgpu_ctx::ARENA_QUOT, (total + 2 * MSM_MAX_BATCH) * sizeof(fr_t), &aerr);
    if (!q_all) return cuda_fail(ctx, aerr, "cudaMalloc(quotient)");
    fr_t *nz = q_all + total, *val = nz + MSM_MAX_BATCH;
    // 1/(z - j) and the span products depend on z only: computed once per call for the longest vector (a shorter vector uses a
    // prefix of them).  Nothing is kept across calls - a repeated proof of the same trace redoes this work.
    fr_t zf; memcpy(zf.l, z->l, 32);
    MsmBasis basis[MSM_MAX_BATCH]; const fr_t* sc[MSM_MAX_BATCH]; size_t n[MSM_MAX_BATCH];
    {
        KernelTimer kt(ctx, "open_bary");
        fr_t nz_max;   // N(z) over the nodes of the longest vector, from the host leg of the batch inversion
        {
            fr_t* inv = (fr_t*)arena_get(ctx, tsgpu_ctx::ARENA_BARY, (mmax + lag_binv_scratch(mmax)) * sizeof(fr_t), &aerr);
            if (!inv) return cuda_fail(ctx, aerr, "cudaMalloc(node inverses)");
            ctx->bary_inv = inv; ctx->bary_spans = inv + mmax; ctx->bary_n = mmax; ctx->bary_z = *z;
            unsigned launches = 0;
            TSG_CUDA(ctx, launch_node_inverses(zf, mmax, inv, (fr_t*)ctx->bary_spans, ctx->host_scratch, &nz_max, ctx->sm_count, ctx->stream, &launches));
            ctx->launches += launches;
            memcpy(ctx->host_scratch + 40, &nz_max, sizeof(fr_t));   // pinned copy for the upload below
        }
        size_t off = 0;
        for (size_t i = 0; i < count; ++i) {
            const size_t m = values[i]->n;
            n[i] = m; sc[i] = q_all + off; basis[i] = MsmBasis{nullptr, 0, nullptr, 0};
            if (m) {
                if (log2_exact(m) < 0) return fail(ctx, TSGPU_E_POLYNOMIAL, "evaluation-basis vectors have power-of-two length");
                int rc = tsgpu_srs_lagrange_prepare(ctx, srs, m);
                if (rc) return rc;
                basis[i] = lagrange_msm_basis(srs, m);
                // N(z) over the nodes 0..m-1: product of whole spans, or (m < span) a fresh small product
                if (m == mmax) TSG_CUDA(ctx, cudaMemcpyAsync(nz + i, ctx->host_scratch + 40, sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
                else if (m % LAG_SPAN == 0) TSG_CUDA(ctx, launch_fr_product((const fr_t*)ctx->bary_spans, lag_num_spans(m), nz + i, ctx->stream));
                else TSG_CUDA(ctx, launch_node_product(zf, m, nz + i, ctx->stream));
                TSG_CUDA(ctx, launch_bary_open(values[i]->d, (const fr_t*)ctx->bary_inv, ifact, m, nz + i, ctx->partials, ctx->ticket, val + i, q_all + off, ctx->sm_count, ctx->stream));
                ctx->launches += 3;
            } else {
                TSG_CUDA(ctx, cudaMemsetAsync(val + i, 0, sizeof(fr_t), ctx->stream));
            }
            off += m;
        }
    }
    TSG_CUDA(ctx, cudaMemcpyAsync(ctx->host_out, val, count * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    int rc = msm_device_batch(ctx, (int)count, basis, sc, n, proofs);
    if (rc) return rc;
    TSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(out_values, ctx->host_out, count * 32);
    return TSGPU_OK;
}
int tsgpu_kzg_open_values_dev(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_poly* values, const tsgpu_fr* z, tsgpu_fr* value, tsgpu_g1* proof) {
    if (!values || !value || !proof) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    return tsgpu_kzg_open_values_batch_dev(ctx, srs, &values, 1, z, value, proof);
}
// ------------------------------------------------------------------------------------------- KZGVectorCommitment (src/commitments.rs:407-483)
// commit(vector) = KZGCommitment::commit(lagrange_interpolate((i, vector[i]))) - any length (no padding: the degree is < len)
int tsgpu_vector_commit(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_fr* vector, size_t n, tsgpu_g1* out) {
    if (!ctx || !srs || !out || (!vector && n)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (n > srs->n) return fail(ctx, TSGPU_E_COMMITMENT, "Polynomial degree exceeds setup size");
    if (n == 0) { G1J id = G1J::identity(); memcpy(out, &id, 96); return TSGPU_OK; }
    tsgpu_poly* p = nullptr;
    int rc = interpolate_any_dev(ctx, vector, n, &p);
    if (rc) return rc;
    rc = tsgpu_kzg_commit_dev(ctx, srs, p, out);
    tsgpu_poly_free(ctx, p);
    return rc;
}
// open(vector, index): (vector[index], KZGCommitment::open(poly, Fr::from(index)).1); "Index out of bounds" beyond the vector
int tsgpu_vector_open(tsgpu_ctx* ctx, const tsgpu_srs* srs, const tsgpu_fr* vector, size_t n, size_t index, tsgpu_fr* value, tsgpu_g1* proof) {
    if (!ctx || !srs || !value || !proof || (!vector && n)) return fail(ctx, TSGPU_E_INVALID_PARAMETERS, "null argument");
    if (index >= n) return fail(ctx, TSGPU_E_COMMITMENT, "Index out of bounds");                       // commitments.rs:436-440
    if (n - 1 > srs->n) return fail(ctx, TSGPU_E_COMMITMENT, "Polynomial degree exceeds setup size");
    tsgpu_poly* p = nullptr;
    int rc = interpolate_any_dev(ctx, vector, n, &p);
    if (rc) return rc;
    Fr64 pt = Fr64::from_u64((uint64_t)index);
    tsgpu_fr z; memcpy(z.l, pt.l, 32);
    tsgpu_fr opened;
    rc = tsgpu_kzg_open_dev(ctx, srs, p, &z, &opened, proof);
    tsgpu_poly_free(ctx, p);
    if (rc) return rc;
    if (memcmp(opened.l, vector[index].l, 32)) return fail(ctx, TSGPU_E_COMMITMENT, "Opened value does not match vector entry");   // commitments.rs:461-465
    *value = vector[index];
    return TSGPU_OK;
}

// group addition of two G1Projective values on the CPU (combining per-rank MSM results)
void tsgpu_g1_add(const tsgpu_g1* a, const tsgpu_g1* b, tsgpu_g1* out) {
    G1J x, y; memcpy(&x, a, 96); memcpy(&y, b, 96);
    G1J r = x.add(y); memcpy(out, &r, 96);
}
}
