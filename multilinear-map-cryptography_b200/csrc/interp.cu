// interp.cu - interpolation on the points 0, 1, ..., n-1 in O(n log^2 n) on the device.
//
// Twist::prove / Shout::prove turn each padded vector into monomial coefficients with
// poly_utils::lagrange_interpolate over x_i = i (src/twist.rs:307-315, src/shout.rs:277-285,
// src/polynomials.rs:301-352: O(n^3) multiplications, n^2 inversions).  The interpolant is unique, so the
// same coefficients are produced here by
//   A. Newton forward differences as one convolution:   c_k = sum_j (-1)^(k-j) v_j / (j! (k-j)!)
//      giving P(x) = sum_k c_k x(x-1)...(x-k+1)
//   B. falling-factorial -> monomial basis, bottom-up:  Q_{2s}(x) = A(x) + F_s(x) B(x - s),
//      F_s(x) = x(x-1)...(x-s+1); blocks of 32 are converted directly, each doubling level is two batched
//      convolutions (Taylor shift of B, product with F_s) over all blocks at once.
// Everything that does not depend on the input (factorials, twiddles, spectra of the shift weights and of
// F_s, scaled by 1/m) is computed once per context and cached in HBM.
#include <map>
#include <vector>
#include "context.cuh"
#include "fr_device.cuh"
#include "interp.cuh"
#include "msm.cuh"
#include "ntt.cuh"
#include "../host/field64.hpp"

namespace tsg {

using host::Fr64;

constexpr unsigned BASE_LOG = 5;
constexpr size_t BASE = (size_t)1 << BASE_LOG;   // blocks converted directly

// ------------------------------------------------------------------------------------------ kernels
// out[i] = in[i] * tbl[i] (negated for odd i when alt), zero for i in [n, total)
__global__ void k_mul_table(fr_t* out, const fr_t* in, const fr_t* tbl, size_t n, size_t total, int alt, int has_in) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        fr_t v = fr_t::zero();
        if (i < n) {
            v = ld256_nc(tbl + i);
            if (has_in) v = v * ld256_nc(in + i);
            if (alt && (i & 1)) v = v.neg();
        }
        st256(out + i, v);
    }
}
// batched: out[b*M + j] = (j < len) ? in[b*in_stride + in_off + len-1-j] * fact[len-1-j] : 0
__global__ void k_rev_mul_fact(fr_t* out, const fr_t* in, size_t in_stride, size_t in_off, size_t len, unsigned logM, const fr_t* fact, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t M = (size_t)1 << logM;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        size_t b = t >> logM, j = t & (M - 1);
        fr_t v = fr_t::zero();
        if (j < len) v = ld256_nc(in + b * in_stride + in_off + (len - 1 - j)) * ld256_nc(fact + (len - 1 - j));
        st256(out + t, v);
    }
}
// data[b*M + i] *= spec[i]
__global__ void k_pointwise(fr_t* data, const fr_t* spec, unsigned logM, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t M = (size_t)1 << logM;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride)
        st256(data + t, ld256(data + t) * ld256_nc(spec + (t & (M - 1))));
}
// out[b*M + k] = (k < len) ? in[b*M + len-1-k] * ifact[k] : 0
__global__ void k_unrev_mul_ifact(fr_t* out, const fr_t* in, size_t len, unsigned logM, const fr_t* ifact, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t M = (size_t)1 << logM;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        size_t b = t >> logM, k = t & (M - 1);
        fr_t v = fr_t::zero();
        if (k < len) v = ld256_nc(in + (b << logM) + (len - 1 - k)) * ld256_nc(ifact + k);
        st256(out + t, v);
    }
}
// blk[b*2s + i] = (i < s) ? blk + U : U
__global__ void k_merge(fr_t* cur, const fr_t* U, unsigned logM, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t M = (size_t)1 << logM, s = M >> 1;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        fr_t u = ld256_nc(U + t);
        if ((t & (M - 1)) < s) u = u + ld256(cur + t);
        st256(cur + t, u);
    }
}
// out[i] = (i < len) ? a^i * ifact[i] : 0   for i < M
__global__ void k_shift_weights(fr_t* out, const fr_t a, size_t len, const fr_t* ifact, size_t M) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t span = 64, nch = (M + span - 1) / span;
    for (size_t ch = (size_t)blockIdx.x * blockDim.x + threadIdx.x; ch < nch; ch += stride) {
        size_t b = ch * span, e = b + span < M ? b + span : M;
        fr_t cur = a.pow_u64(b);
        for (size_t i = b; i < e; ++i) {
            st256(out + i, i < len ? cur * ld256_nc(ifact + i) : fr_t::zero());
            cur = cur * a;
        }
    }
}
// data[i] *= scale
__global__ void k_scale(fr_t* data, const fr_t scale, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) st256(data + i, ld256(data + i) * scale);
}
// falling-factorial -> monomial for independent blocks of `blen` <= 32 coefficients (one thread per block):
// P = c_0 + x (c_1 + (x-1)(c_2 + ...)), expanded by Horner in the monomial basis.  small[k] = Fr(k).
__global__ void __launch_bounds__(32) k_ff2mono_base(fr_t* cur, size_t nblocks, unsigned blen, const fr_t* small) {
    __shared__ fr_t sh[32 * 33];
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    fr_t* m = sh + threadIdx.x * 33;           // 33-element stride: conflict-free across threads
    fr_t* blk = cur + b * blen;
    for (unsigned i = 0; i < blen; ++i) m[i] = fr_t::zero();
    m[0] = ld256(blk + blen - 1);
    unsigned deg = 0;
    for (unsigned k = blen - 1; k-- > 0;) {
        // m <- m * (x - k) + c_k
        fr_t kk = ld256_nc(small + k);
        ++deg;
        for (unsigned t = deg; t >= 1; --t) m[t] = m[t - 1] - (k ? kk * m[t] : fr_t::zero());
        m[0] = ld256(blk + k) - (k ? kk * m[0] : fr_t::zero());
    }
    for (unsigned i = 0; i < blen; ++i) st256(blk + i, m[i]);
}

static inline int gridfor(size_t work, int threads, size_t cap) {
    size_t g = (work + threads - 1) / threads;
    if (g < 1) g = 1;
    return (int)(g < cap ? g : cap);
}

// ------------------------------------------------------------------------------------------ plan
struct InterpPlan {
    tsgpu_ctx* ctx;
    unsigned log_max = 0;                 // factorial tables cover 0 .. 2^log_max
    fr_t *fact = nullptr, *ifact = nullptr, *small = nullptr;
    std::map<unsigned, fr_t*> tw, twi;    // by logm
    std::map<unsigned, fr_t*> what, fhat; // by log s: spectra (size 2s), pre-scaled by 1/(2s)
    std::map<unsigned, fr_t*> bhat;       // by log n: spectrum of (-1)^j / j! (size 2n), pre-scaled by 1/(2n)
    std::map<unsigned, fr_t*> fs;         // by log s: F_s coefficients (s + 1)
    Fr64 root28;
    unsigned launches = 0;

    explicit InterpPlan(tsgpu_ctx* c) : ctx(c) {
        // omega_{2^28} = 5^((r-1)/2^28)
        uint64_t e[4];
        for (int i = 0; i < 4; ++i) e[i] = Fr64::modl(i);
        e[0] -= 1;
        for (int i = 0; i < 4; ++i) e[i] = (e[i] >> 28) | (i < 3 ? e[i + 1] << 36 : 0);
        Fr64 g = Fr64::from_u64(5), acc = Fr64::one();
        for (int i = 255; i >= 0; --i) { acc = acc.sqr(); if ((e[i >> 6] >> (i & 63)) & 1) acc = acc * g; }
        root28 = acc;
    }
    ~InterpPlan() {
        cudaFree(fact); cudaFree(ifact); cudaFree(small);
        for (auto* m : {&tw, &twi, &what, &fhat, &bhat, &fs}) for (auto& kv : *m) cudaFree(kv.second);
    }
    cudaStream_t st() const { return ctx->stream; }
    int sms() const { return ctx->sm_count; }
    size_t cap() const { return (size_t)ctx->sm_count * 8; }

    cudaError_t ensure_factorials(unsigned logn) {
        if (fact && logn <= log_max) return cudaSuccess;
        cudaFree(fact); cudaFree(ifact); fact = ifact = nullptr;
        const size_t n = (size_t)1 << logn;
        std::vector<Fr64> f(n + 1), fi(n + 1);
        f[0] = Fr64::one();
        for (size_t i = 1; i <= n; ++i) f[i] = f[i - 1] * Fr64::from_u64(i);
        fi[n] = f[n].inverse();
        for (size_t i = n; i > 0; --i) fi[i - 1] = fi[i] * Fr64::from_u64(i);
        cudaError_t e;
        if ((e = cudaMalloc((void**)&fact, (n + 1) * 32))) return e;
        if ((e = cudaMalloc((void**)&ifact, (n + 1) * 32))) return e;
        if ((e = cudaMemcpy(fact, f.data(), (n + 1) * 32, cudaMemcpyHostToDevice))) return e;
        if ((e = cudaMemcpy(ifact, fi.data(), (n + 1) * 32, cudaMemcpyHostToDevice))) return e;
        if (!small) {
            std::vector<Fr64> sm(64);
            for (size_t i = 0; i < 64; ++i) sm[i] = Fr64::from_u64(i);
            if ((e = cudaMalloc((void**)&small, 64 * 32))) return e;
            if ((e = cudaMemcpy(small, sm.data(), 64 * 32, cudaMemcpyHostToDevice))) return e;
        }
        log_max = logn;
        return cudaSuccess;
    }
    cudaError_t twiddles(unsigned logm, fr_t** f, fr_t** inv) {
        if (!tw.count(logm)) {
            const size_t half = logm ? (size_t)1 << (logm - 1) : 1;
            Fr64 w = root28;
            for (unsigned i = logm; i < 28; ++i) w = w.sqr();
            Fr64 wi = w.inverse();
            fr_t *a = nullptr, *b = nullptr;
            cudaError_t e;
            if ((e = cudaMalloc((void**)&a, half * 32))) return e;
            if ((e = cudaMalloc((void**)&b, half * 32))) return e;
            fr_t wf, wif; memcpy(wf.l, w.l, 32); memcpy(wif.l, wi.l, 32);
            if ((e = launch_tau_powers(wf, 0, half, a, sms(), st()))) return e;
            if ((e = launch_tau_powers(wif, 0, half, b, sms(), st()))) return e;
            launches += 2;
            tw[logm] = a; twi[logm] = b;
        }
        *f = tw[logm]; *inv = twi[logm];
        return cudaSuccess;
    }
    static fr_t to_fr(const Fr64& x) { fr_t r; memcpy(r.l, x.l, 32); return r; }
    static fr_t inv_pow2(unsigned logm) { return to_fr(Fr64::from_u64((uint64_t)1 << logm).inverse()); }
    static fr_t neg_small(uint64_t s) { return to_fr(Fr64::from_u64(s).neg()); }

    // cyclic convolution helpers on buffers of size M = 2^logM (batch arrays): data <- IFFT(FFT(data) * spec)
    cudaError_t conv_with_spectrum(fr_t* data, unsigned logM, size_t batch, const fr_t* spec) {
        fr_t *f, *inv; cudaError_t e;
        if ((e = twiddles(logM, &f, &inv))) return e;
        if ((e = ntt_forward(data, logM, batch, f, sms(), st(), &launches))) return e;
        const size_t total = batch << logM;
        k_pointwise<<<gridfor(total, 256, cap()), 256, 0, st()>>>(data, spec, logM, total); ++launches;
        return ntt_inverse_unscaled(data, logM, batch, inv, sms(), st(), &launches);
    }
    // spectrum (scaled by 1/M) of the Taylor-shift weights a^i / i!, i < len
    cudaError_t shift_spectrum(const fr_t& a, size_t len, unsigned logM, fr_t** out) {
        const size_t M = (size_t)1 << logM;
        fr_t* buf; cudaError_t e;
        if ((e = cudaMalloc((void**)&buf, M * 32))) return e;
        k_shift_weights<<<gridfor((M + 63) / 64, 128, cap()), 128, 0, st()>>>(buf, a, len, ifact, M); ++launches;
        fr_t *f, *inv;
        if ((e = twiddles(logM, &f, &inv))) return e;
        if ((e = ntt_forward(buf, logM, 1, f, sms(), st(), &launches))) return e;
        k_scale<<<gridfor(M, 256, cap()), 256, 0, st()>>>(buf, inv_pow2(logM), M); ++launches;
        *out = buf;
        return cudaGetLastError();
    }
    // F_s coefficient vectors (s + 1 entries) for log s = BASE_LOG .. upto
    cudaError_t ensure_fs(unsigned upto) {
        cudaError_t e;
        if (!fs.count(BASE_LOG)) {
            std::vector<Fr64> p(BASE + 1, Fr64::zero());
            p[0] = Fr64::one();
            size_t deg = 0;
            for (size_t i = 0; i < BASE; ++i) {   // multiply by (x - i)
                Fr64 c = Fr64::from_u64(i);
                ++deg;
                for (size_t t = deg; t >= 1; --t) p[t] = p[t - 1] - c * p[t];
                p[0] = Fr64::zero() - c * p[0];
            }
            fr_t* d;
            if ((e = cudaMalloc((void**)&d, (BASE + 1) * 32))) return e;
            if ((e = cudaMemcpy(d, p.data(), (BASE + 1) * 32, cudaMemcpyHostToDevice))) return e;
            fs[BASE_LOG] = d;
        }
        for (unsigned ls = BASE_LOG; ls < upto; ++ls) {
            if (fs.count(ls + 1)) continue;
            // F_{2s}(x) = F_s(x) * F_s(x - s): Taylor shift (len s+1) and product, both in size M = 4s
            const size_t s = (size_t)1 << ls, len = s + 1;
            const unsigned logM = ls + 2; const size_t M = (size_t)1 << logM;
            fr_t *U, *V, *wspec;
            if ((e = cudaMalloc((void**)&U, M * 32))) return e;
            if ((e = cudaMalloc((void**)&V, M * 32))) return e;
            if ((e = shift_spectrum(neg_small(s), len, logM, &wspec))) return e;
            k_rev_mul_fact<<<gridfor(M, 256, cap()), 256, 0, st()>>>(U, fs[ls], 0, 0, len, logM, fact, M); ++launches;
            if ((e = conv_with_spectrum(U, logM, 1, wspec))) return e;
            k_unrev_mul_ifact<<<gridfor(M, 256, cap()), 256, 0, st()>>>(V, U, len, logM, ifact, M); ++launches;   // V = F_s(x - s), padded
            // spectrum of F_s (scaled) then product
            fr_t *f, *inv;
            if ((e = twiddles(logM, &f, &inv))) return e;
            if ((e = cudaMemsetAsync(U, 0, M * 32, st()))) return e;
            if ((e = cudaMemcpyAsync(U, fs[ls], len * 32, cudaMemcpyDeviceToDevice, st()))) return e;
            if ((e = ntt_forward(U, logM, 1, f, sms(), st(), &launches))) return e;
            k_scale<<<gridfor(M, 256, cap()), 256, 0, st()>>>(U, inv_pow2(logM), M); ++launches;
            if ((e = conv_with_spectrum(V, logM, 1, U))) return e;
            fr_t* d;
            if ((e = cudaMalloc((void**)&d, (2 * s + 1) * 32))) return e;
            if ((e = cudaMemcpyAsync(d, V, (2 * s + 1) * 32, cudaMemcpyDeviceToDevice, st()))) return e;
            if ((e = cudaStreamSynchronize(st()))) return e;
            cudaFree(U); cudaFree(V); cudaFree(wspec);
            fs[ls + 1] = d;
        }
        return cudaSuccess;
    }
    // per-level spectra for merging blocks of size s into 2s
    cudaError_t ensure_level(unsigned ls) {
        if (what.count(ls)) return cudaSuccess;
        cudaError_t e;
        if ((e = ensure_fs(ls))) return e;
        const size_t s = (size_t)1 << ls; const unsigned logM = ls + 1; const size_t M = 2 * s;
        fr_t* w;
        if ((e = shift_spectrum(neg_small(s), s, logM, &w))) return e;
        fr_t* fh;
        if ((e = cudaMalloc((void**)&fh, M * 32))) return e;
        if ((e = cudaMemsetAsync(fh, 0, M * 32, st()))) return e;
        // F_s has s + 1 coefficients (degree s < 2s): no wrap-around
        if ((e = cudaMemcpyAsync(fh, fs[ls], (s + 1) * 32, cudaMemcpyDeviceToDevice, st()))) return e;
        fr_t *f, *inv;
        if ((e = twiddles(logM, &f, &inv))) return e;
        if ((e = ntt_forward(fh, logM, 1, f, sms(), st(), &launches))) return e;
        k_scale<<<gridfor(M, 256, cap()), 256, 0, st()>>>(fh, inv_pow2(logM), M); ++launches;
        what[ls] = w; fhat[ls] = fh;
        return cudaGetLastError();
    }
    cudaError_t ensure_bhat(unsigned logn) {
        if (bhat.count(logn)) return cudaSuccess;
        const size_t n = (size_t)1 << logn; const unsigned logM = logn + 1; const size_t M = 2 * n;
        fr_t* b; cudaError_t e;
        if ((e = cudaMalloc((void**)&b, M * 32))) return e;
        k_mul_table<<<gridfor(M, 256, cap()), 256, 0, st()>>>(b, nullptr, ifact, n, M, 1, 0); ++launches;
        fr_t *f, *inv;
        if ((e = twiddles(logM, &f, &inv))) return e;
        if ((e = ntt_forward(b, logM, 1, f, sms(), st(), &launches))) return e;
        k_scale<<<gridfor(M, 256, cap()), 256, 0, st()>>>(b, inv_pow2(logM), M); ++launches;
        bhat[logn] = b;
        return cudaGetLastError();
    }
    cudaError_t prepare(unsigned logn) {
        cudaError_t e;
        if ((e = ensure_factorials(logn > log_max ? logn : log_max))) return e;
        if (logn > BASE_LOG) {
            if ((e = ensure_bhat(logn))) return e;
            for (unsigned ls = BASE_LOG; ls < logn; ++ls) if ((e = ensure_level(ls))) return e;
        }
        return cudaSuccess;
    }

    // vals (n = 2^logn entries, natural order) -> coeffs (n entries, low -> high).  vals and coeffs may alias.
    // n_valid < n: only the first n_valid values are interpolation data (degree < n_valid); the Newton coefficients beyond are zero
    cudaError_t run(const fr_t* vals, unsigned logn, fr_t* coeffs, size_t n_valid) {
        cudaError_t e;
        const size_t n = (size_t)1 << logn;
        if ((e = prepare(logn))) return e;
        if (logn <= BASE_LOG) {
            // tiny: Newton coefficients by direct O(n^2) differences are not worth a kernel; reuse the base kernel
            // on c_k computed with the same convolution at size 2n would need bhat; do it on the host path of step A:
            // the base kernel converts falling-factorial -> monomial, step A below handles any n >= 1.
        }
        // ---- step A: c = (v / j!) * ((-1)^j / j!)  truncated to n
        fr_t *bufA;
        const unsigned logM = logn + 1; const size_t M = 2 * n;
        bufA = (fr_t*)arena_get(ctx, tsgpu_ctx::ARENA_INTERP, M * 32, &e);
        if (!bufA) return e;
        if ((e = ensure_bhat(logn))) return e;
        k_mul_table<<<gridfor(M, 256, cap()), 256, 0, st()>>>(bufA, vals, ifact, n, M, 0, 1); ++launches;
        if ((e = conv_with_spectrum(bufA, logM, 1, bhat[logn]))) return e;
        if ((e = cudaMemcpyAsync(coeffs, bufA, n * 32, cudaMemcpyDeviceToDevice, st()))) return e;
        if (n_valid < n && (e = cudaMemsetAsync(coeffs + n_valid, 0, (n - n_valid) * 32, st()))) return e;
        // ---- step B base: blocks of min(n, 32)
        const unsigned blen_log = logn < BASE_LOG ? logn : BASE_LOG;
        const size_t blen = (size_t)1 << blen_log, nblk = n >> blen_log;
        k_ff2mono_base<<<(unsigned)((nblk + 31) / 32), 32, 0, st()>>>(coeffs, nblk, (unsigned)blen, small); ++launches;
        // ---- step B levels
        fr_t* U = bufA;            // n elements
        fr_t* V = bufA + n;        // n elements
        for (unsigned ls = BASE_LOG; ls < logn; ++ls) {
            const size_t s = (size_t)1 << ls; const unsigned lM = ls + 1; const size_t nblocks = n >> lM;
            // Taylor shift of every block's upper half B(x) -> B(x - s)
            k_rev_mul_fact<<<gridfor(n, 256, cap()), 256, 0, st()>>>(U, coeffs, 2 * s, s, s, lM, fact, n); ++launches;
            if ((e = conv_with_spectrum(U, lM, nblocks, what[ls]))) return e;
            k_unrev_mul_ifact<<<gridfor(n, 256, cap()), 256, 0, st()>>>(V, U, s, lM, ifact, n); ++launches;
            // times F_s
            if ((e = conv_with_spectrum(V, lM, nblocks, fhat[ls]))) return e;
            k_merge<<<gridfor(n, 256, cap()), 256, 0, st()>>>(coeffs, V, lM, n); ++launches;
        }
        return cudaGetLastError();
    }
};

InterpPlan* interp_plan(tsgpu_ctx* ctx) {
    if (!ctx->interp) ctx->interp = new InterpPlan(ctx);
    return (InterpPlan*)ctx->interp;
}
void interp_destroy(tsgpu_ctx* ctx) {
    delete (InterpPlan*)ctx->interp;
    ctx->interp = nullptr;
}
cudaError_t interp_prepare(tsgpu_ctx* ctx, unsigned logn) {
    InterpPlan* p = interp_plan(ctx);
    p->launches = 0;
    cudaError_t e = p->prepare(logn);
    ctx->launches += p->launches;
    return e;
}
cudaError_t interp_factorials(tsgpu_ctx* ctx, unsigned logn, const fr_t** ifact) {
    InterpPlan* p = interp_plan(ctx);
    cudaError_t e = p->ensure_factorials(logn > p->log_max ? logn : p->log_max);
    *ifact = p->ifact;
    return e;
}
cudaError_t interp_run(tsgpu_ctx* ctx, const fr_t* vals, unsigned logn, fr_t* coeffs, size_t n_valid) {
    InterpPlan* p = interp_plan(ctx);
    p->launches = 0;
    cudaError_t e = p->run(vals, logn, coeffs, n_valid ? n_valid : (size_t)1 << logn);
    ctx->launches += p->launches;
    return e;
}

}  // namespace tsg
