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

namespace tsg {

static inline int gridfor(size_t work, int threads, size_t cap) {
    size_t g = (work + threads - 1) / threads;
    if (g < 1) g = 1;
    return (int)(g < cap ? g : cap);
}

// inv[j] = 1 / (pt - j) for j < n (pt must not be one of 0..n-1);  span_prod[ch] = prod_{j in span ch} (pt - j)
__global__ void __launch_bounds__(128) k_node_inverses(const fr_t pt, size_t n, fr_t* inv, fr_t* span_prod) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t nch = (n + LAG_SPAN - 1) / LAG_SPAN;
    const fr_t one = fr_t::one();
    for (size_t ch = (size_t)blockIdx.x * blockDim.x + threadIdx.x; ch < nch; ch += stride) {
        const size_t b = ch * LAG_SPAN, e = b + LAG_SPAN < n ? b + LAG_SPAN : n;
        fr_t d = pt - fr_t::from_u64(b);          // pt - j, stepped by -1
        fr_t acc = one;
        for (size_t j = b; j < e; ++j) { st256(inv + j, acc); acc = acc * d; d = d - one; }
        st256(span_prod + ch, acc);
        fr_t ia = acc.inverse();
        for (size_t j = e; j-- > b;) {
            d = d + one;                          // back to pt - j
            fr_t pref = ld256(inv + j);
            st256(inv + j, ia * pref);
            ia = ia * d;
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
__global__ void __launch_bounds__(256) k_lagrange_scalars(const fr_t* inv, const fr_t* ifact, const fr_t* ntau, size_t n, fr_t* scal) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const fr_t N = *ntau;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
        st256(scal + j, node_weight(ifact, j, n) * ld256_nc(inv + j) * N);
}

// out[0] = N(z) * sum_j w_j v_j / (z - j)
struct BaryEpilogue {
    const fr_t* nz; fr_t* out;
    __device__ void operator()(fr_t (&v)[1]) const { *out = v[0] * *nz; }
};
__global__ void __launch_bounds__(256) k_bary_sum(const fr_t* vals, const fr_t* inv, const fr_t* ifact, size_t n, const fr_t* nz,
                                                  fr_t* partials, unsigned int* ticket, fr_t* out) {
    __shared__ fr_t smem[32];
    wide_acc<FrP> acc; acc.clear();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        fr_t v = ld256_nc(vals + j);
        if (v.is_zero()) continue;                 // zero padding and never-written cells
        acc.add_product(v * ld256_nc(inv + j), node_weight(ifact, j, n));
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

cudaError_t launch_node_inverses(const fr_t& pt, size_t n, fr_t* inv, fr_t* span_prod, int sm_count, cudaStream_t s) {
    const size_t nch = lag_num_spans(n);
    k_node_inverses<<<gridfor(nch, 128, (size_t)sm_count * 16), 128, 0, s>>>(pt, n, inv, span_prod);
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
cudaError_t launch_lagrange_scalars(const fr_t* inv, const fr_t* ifact, const fr_t* ntau, size_t n, fr_t* scal, int sm_count, cudaStream_t s) {
    k_lagrange_scalars<<<gridfor(n, 256, (size_t)sm_count * 8), 256, 0, s>>>(inv, ifact, ntau, n, scal);
    return cudaGetLastError();
}
cudaError_t launch_bary_open(const fr_t* vals, const fr_t* inv, const fr_t* ifact, size_t n, const fr_t* nz, fr_t* partials, unsigned int* ticket,
                             fr_t* value, fr_t* q, int sm_count, cudaStream_t s) {
    k_bary_sum<<<gridfor(n, 256, (size_t)sm_count * 2), 256, 0, s>>>(vals, inv, ifact, n, nz, partials, ticket, value);
    k_bary_quotient<<<gridfor(n, 256, (size_t)sm_count * 8), 256, 0, s>>>(vals, inv, value, n, q);
    return cudaGetLastError();
}

}  // namespace tsg
