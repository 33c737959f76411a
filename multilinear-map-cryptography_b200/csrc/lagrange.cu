// lagrange.cu - the evaluation-basis ("Lagrange") form of the KZG steps of Twist::prove / Shout::prove.
//
// The reference commits to a padded vector v_0..v_{n-1} by first interpolating it on x_i = i
// (vector_to_polynomial -> poly_utils::lagrange_interpolate, src/twist.rs:307-315, src/polynomials.rs:301-352)
// and then summing coeff_i * [tau^i]_1 (KZGCommitment::commit, src/commitments.rs:162-180).  The group
// element is a linear function of the VALUES:  commit(P) = sum_j v_j * [L_j(tau)]_1  with the Lagrange basis
// L_j of the nodes 0..n-1.  With the basis points [L_j(tau)]_1 prepared once per SRS the commitment is one
// MSM over the raw values (16-bit addresses, 64-bit memory values: few non-zero windows) and the
// interpolation leaves the proving path.  The opening (KZGCommitment::open, src/commitments.rs:182-199:
// Horner value, (P - v) / (x - z), commit) becomes
//     v      = P(z) = N(z) * sum_j w_j v_j / (z - j),   N(z) = prod_k (z - k),   w_j = (-1)^(n-1-j) / (j! (n-1-j)!)
//     Q(j)   = (v_j - v) / (j - z)          (the quotient has degree n - 2, so its n values determine it)
//     proof  = sum_j Q(j) * [L_j(tau)]_1
// Same field element v, same group elements: proofs stay byte-identical (tests/test_gpu_lagrange.py).
// The kernels below are streaming passes with a Montgomery batch inversion per thread.
#include "fr_device.cuh"
#include "lagrange.cuh"
#include <cstring>
#include <vector>
#include "../host/field64.hpp"

namespace tsg {

static inline int gridfor(size_t work, int threads, size_t cap) {
    size_t g = (work + threads - 1) / threads;
    if (g < 1) g = 1;
    return (int)(g < cap ? g : cap);
}

// ---- batch inversion of d_j = pt - j (j < n) with ONE field inversion, done on the host ------------------------------------
// A Fermat inversion is a chain of ~380 dependent products (~0.13 ms for a lone thread), so the classic per-thread Montgomery
// trick is latency-bound.  Here the prefix-product / back-substitution passes are applied level by level: level 0 over the
// d_j in spans of LAG_SPAN, level 1 over the span products, ... until at most LAG_SPAN products remain; those go to the host
// (1 KiB), which inverts them with its native 64-bit arithmetic (microseconds) and sends the inverses back; the levels are
// then unwound.  Every launch is at most 2 x LAG_SPAN products deep.
// up:   pre[j] = product of the elements before j in its span; span_prod[ch] = product of span ch
template <bool GEN>
__global__ void __launch_bounds__(128) k_binv_up(const fr_t pt, const fr_t* elems, size_t n, fr_t* pre, fr_t* span_prod) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t nch = (n + LAG_SPAN - 1) / LAG_SPAN;
    const fr_t one = fr_t::one();
    for (size_t ch = (size_t)blockIdx.x * blockDim.x + threadIdx.x; ch < nch; ch += stride) {
        const size_t b = ch * LAG_SPAN, e = b + LAG_SPAN < n ? b + LAG_SPAN : n;
        fr_t d = GEN ? pt - fr_t::from_u64(b) : fr_t::zero();
        fr_t acc = one;
        for (size_t j = b; j < e; ++j) {
            st256(pre + j, acc);
            if (GEN) { acc = acc * d; d = d - one; } else acc = acc * ld256(elems + j);
        }
        st256(span_prod + ch, acc);
    }
}
// down: span_inv[ch] = 1 / (product of span ch); pre[j] (prefix products) is replaced by 1 / element_j
template <bool GEN>
__global__ void __launch_bounds__(128) k_binv_down(const fr_t pt, const fr_t* elems, size_t n, fr_t* pre, const fr_t* span_inv) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t nch = (n + LAG_SPAN - 1) / LAG_SPAN;
    const fr_t one = fr_t::one();
    for (size_t ch = (size_t)blockIdx.x * blockDim.x + threadIdx.x; ch < nch; ch += stride) {
        const size_t b = ch * LAG_SPAN, e = b + LAG_SPAN < n ? b + LAG_SPAN : n;
        fr_t ia = ld256(span_inv + ch);
        fr_t d = GEN ? pt - fr_t::from_u64(e) : fr_t::zero();   // stepped up to pt - j below
        for (size_t j = e; j-- > b;) {
            fr_t el;
            if (GEN) { d = d + one; el = d; } else el = ld256(elems + j);
            fr_t p = ld256(pre + j);
            st256(pre + j, ia * p);
            ia = ia * el;
        }
    }
}

// out[0] = scale * prod_i in[i]   (one block)
__global__ void __launch_bounds__(LAG_PROD_THREADS) k_fr_product(const fr_t* in, size_t count, fr_t* out) {
    __shared__ fr_t sh[LAG_PROD_THREADS];
    fr_t acc = fr_t::one();
    for (size_t i = threadIdx.x; i < count; i += blockDim.x) acc = acc * ld256(in + i);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (unsigned s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] = sh[threadIdx.x] * sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

// barycentric weight of node j among 0..n-1 from the inverse-factorial table
__device__ __forceinline__ fr_t node_weight(const fr_t* ifact, size_t j, size_t n) {
    fr_t w = ld256_nc(ifact + j) * ld256_nc(ifact + (n - 1 - j));
    return ((n - 1 - j) & 1) ? w.neg() : w;
}

// scal[j] = L_j(tau) = N(tau) w_j / (tau - j);  inv holds 1/(tau - j) on entry, *ntau = N(tau)
// slice form: inv / scal hold the `count` nodes first .. first + count - 1 of the n-node domain
__global__ void __launch_bounds__(256) k_lagrange_scalars(const fr_t* inv, const fr_t* ifact, const fr_t* ntau, size_t n, size_t first, size_t count, fr_t* scal) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const fr_t N = *ntau;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride)
        st256(scal + j, node_weight(ifact, first + j, n) * ld256_nc(inv + j) * N);
}

// out[0] = N(z) * sum_j w_j v_j / (z - j)
struct BaryEpilogue {
    const fr_t* nz; fr_t* out;
    __device__ void operator()(fr_t (&v)[1]) const { *out = v[0] * *nz; }
};
// vals / inv hold the `count` nodes first .. first + count - 1 of the n-node domain (the whole domain: first = 0, count = n)
__global__ void __launch_bounds__(256) k_bary_sum(const fr_t* vals, const fr_t* inv, const fr_t* ifact, size_t n, size_t first, size_t count, const fr_t* nz,
                                                  fr_t* partials, unsigned int* ticket, fr_t* out) {
    __shared__ fr_t smem[32];
    wide_acc<FrP> acc; acc.clear();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride) {
        fr_t v = ld256_nc(vals + j);
        if (v.is_zero()) continue;                 // zero padding and never-written cells
        acc.add_product(v * ld256_nc(inv + j), node_weight(ifact, first + j, n));
    }
    fr_t v[1]; v[0] = acc.reduce();
    grid_finish_sum<fr_t, 1>(v, partials, ticket, smem, BaryEpilogue{nz, out});
}

// q[j] = (value - v_j) / (z - j)  = Q(j)
__global__ void __launch_bounds__(256) k_bary_quotient(const fr_t* vals, const fr_t* inv, const fr_t* value, size_t n, fr_t* q) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const fr_t v = *value;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
        st256(q + j, (v - ld256_nc(vals + j)) * ld256_nc(inv + j));
}

// scratch: lag_binv_scratch(n) elements.  inv[j] = 1/(pt - j); span_prod (level-0 span products) is left at scratch[0 .. lag_num_spans(n));
// *total = prod_j (pt - j) on the HOST.  Synchronises the stream once (the host inversion sits in the middle).
size_t lag_binv_scratch(size_t n) {
    size_t tot = 0, m = n;
    do { m = lag_num_spans(m); tot += 2 * m; } while (m > LAG_SPAN);
    return tot + 2 * LAG_SPAN;
}
cudaError_t launch_node_inverses(const fr_t& pt, size_t n, fr_t* inv, fr_t* scratch, fr_t* host_pinned, fr_t* total_host, int sm_count, cudaStream_t s,
                                 unsigned* launches) {
    // level sizes: n -> spans(n) -> ... -> top (<= LAG_SPAN elements)
    std::vector<size_t> size{n};
    while (size.back() > LAG_SPAN) size.push_back(lag_num_spans(size.back()));
    const size_t levels = size.size() - 1;     // number of up passes that produce a further device level
    // layout: prod[l] (level l+1 elements = span products of level l) and pre[l] (prefix scratch of level l+1), l = 0 .. levels-1
    std::vector<fr_t*> prod(levels + 1), pre(levels + 1);
    fr_t* p = scratch;
    for (size_t l = 0; l < levels; ++l) { prod[l] = p; p += size[l + 1]; }
    for (size_t l = 1; l <= levels; ++l) { pre[l] = p; p += size[l]; }
    const size_t cap = (size_t)sm_count * 16;
    // up
    for (size_t l = 0; l < levels; ++l) {
        const size_t nch = size[l + 1];
        if (l == 0) k_binv_up<true><<<gridfor(nch, 128, cap), 128, 0, s>>>(pt, nullptr, size[0], inv, prod[0]);
        else k_binv_up<false><<<gridfor(nch, 128, cap), 128, 0, s>>>(pt, prod[l - 1], size[l], pre[l], prod[l]);
        if (launches) ++*launches;
    }
    // top level on the host: elements = prod[levels-1] (or, when n <= LAG_SPAN, the d_j themselves)
    const size_t top = size[levels];
    cudaError_t e;
    if (levels == 0) {
        // tiny n: one span; run the up pass to get prefixes, invert the single product on the host
        k_binv_up<true><<<1, 128, 0, s>>>(pt, nullptr, n, inv, scratch);
        if (launches) ++*launches;
        if ((e = cudaMemcpyAsync(host_pinned, scratch, sizeof(fr_t), cudaMemcpyDeviceToHost, s))) return e;
        if ((e = cudaStreamSynchronize(s))) return e;
        host::Fr64 t = host::Fr64::from_raw((const uint64_t*)host_pinned[0].l);
        memcpy(total_host->l, t.l, 32);
        host::Fr64 ti = t.inverse();
        memcpy(host_pinned[1].l, ti.l, 32);
        if ((e = cudaMemcpyAsync(scratch + 1, host_pinned + 1, sizeof(fr_t), cudaMemcpyHostToDevice, s))) return e;
        k_binv_down<true><<<1, 128, 0, s>>>(pt, nullptr, n, inv, scratch + 1);
        if (launches) ++*launches;
        return cudaGetLastError();
    }
    if ((e = cudaMemcpyAsync(host_pinned, prod[levels - 1], top * sizeof(fr_t), cudaMemcpyDeviceToHost, s))) return e;
    if ((e = cudaStreamSynchronize(s))) return e;
    {
        std::vector<host::Fr64> el(top), prefix(top);
        host::Fr64 acc = host::Fr64::one();
        for (size_t i = 0; i < top; ++i) { el[i] = host::Fr64::from_raw((const uint64_t*)host_pinned[i].l); prefix[i] = acc; acc = acc * el[i]; }
        memcpy(total_host->l, acc.l, 32);
        host::Fr64 ia = acc.inverse();
        for (size_t i = top; i-- > 0;) { host::Fr64 r = ia * prefix[i]; ia = ia * el[i]; memcpy(host_pinned[i].l, r.l, 32); }
    }
    fr_t* top_inv = p;   // top <= LAG_SPAN elements
    if ((e = cudaMemcpyAsync(top_inv, host_pinned, top * sizeof(fr_t), cudaMemcpyHostToDevice, s))) return e;
    // down: level l elements get their inverses from the inverses of level l+1
    const fr_t* upper_inv = top_inv;
    for (size_t l = levels; l-- > 0;) {
        const size_t nch = size[l + 1];
        if (l == 0) k_binv_down<true><<<gridfor(nch, 128, cap), 128, 0, s>>>(pt, nullptr, size[0], inv, upper_inv);
        else { k_binv_down<false><<<gridfor(nch, 128, cap), 128, 0, s>>>(pt, prod[l - 1], size[l], pre[l], upper_inv); upper_inv = pre[l]; }
        if (launches) ++*launches;
    }
    return cudaGetLastError();
}
// out[0] = prod_{j < m} (pt - j) for a short node range (m below one span)
__global__ void k_node_product(const fr_t pt, size_t m, fr_t* out) {
    fr_t acc = fr_t::one(), d = pt;
    const fr_t one = fr_t::one();
    for (size_t j = 0; j < m; ++j) { acc = acc * d; d = d - one; }
    out[0] = acc;
}
cudaError_t launch_node_product(const fr_t& pt, size_t m, fr_t* out, cudaStream_t s) {
    k_node_product<<<1, 1, 0, s>>>(pt, m, out);
    return cudaGetLastError();
}
cudaError_t launch_fr_product(const fr_t* in, size_t count, fr_t* out, cudaStream_t s) {
    k_fr_product<<<1, LAG_PROD_THREADS, 0, s>>>(in, count, out);
    return cudaGetLastError();
}
cudaError_t launch_lagrange_scalars(const fr_t* inv, const fr_t* ifact, const fr_t* ntau, size_t n, size_t first, size_t count, fr_t* scal, int sm_count, cudaStream_t s) {
    k_lagrange_scalars<<<gridfor(count, 256, (size_t)sm_count * 8), 256, 0, s>>>(inv, ifact, ntau, n, first, count, scal);
    return cudaGetLastError();
}
cudaError_t launch_bary_open(const fr_t* vals, const fr_t* inv, const fr_t* ifact, size_t n, const fr_t* nz, fr_t* partials, unsigned int* ticket,
                             fr_t* value, fr_t* q, int sm_count, cudaStream_t s) {
    k_bary_sum<<<gridfor(n, 256, (size_t)sm_count * 2), 256, 0, s>>>(vals, inv, ifact, n, 0, n, nz, partials, ticket, value);
    k_bary_quotient<<<gridfor(n, 256, (size_t)sm_count * 8), 256, 0, s>>>(vals, inv, value, n, q);
    return cudaGetLastError();
}
// sharded opening, first half: *partial = scale * sum over this rank's nodes of w_j v_j / (z - j)   (scale = 1 for a raw partial sum)
cudaError_t launch_bary_partial(const fr_t* vals, const fr_t* inv, const fr_t* ifact, size_t n, size_t first, size_t count, const fr_t* scale,
                                fr_t* partials, unsigned int* ticket, fr_t* partial, int sm_count, cudaStream_t s) {
    k_bary_sum<<<gridfor(count, 256, (size_t)sm_count * 2), 256, 0, s>>>(vals, inv, ifact, n, first, count, scale, partials, ticket, partial);
    return cudaGetLastError();
}
// second half: q[j] = (*value - v_j) / (z - j) over this rank's nodes
cudaError_t launch_bary_quotient(const fr_t* vals, const fr_t* inv, const fr_t* value, size_t count, fr_t* q, int sm_count, cudaStream_t s) {
    k_bary_quotient<<<gridfor(count, 256, (size_t)sm_count * 8), 256, 0, s>>>(vals, inv, value, count, q);
    return cudaGetLastError();
}

}  // namespace tsg
