// lagrange.cuh - launch interface of the evaluation-basis KZG kernels (lagrange.cu)
#pragma once
#include <cuda_runtime.h>
#include "fp.cuh"

namespace tsg {

constexpr size_t LAG_SPAN = 32;            // nodes per batch inversion (one Fermat inverse per span)
constexpr int LAG_PROD_THREADS = 512;
inline size_t lag_num_spans(size_t n) { return (n + LAG_SPAN - 1) / LAG_SPAN; }

// inv[j] = 1/(pt - j), j < n, by a multi-level batch inversion whose single field inversion runs on the host (synchronises the stream once).
// scratch: lag_binv_scratch(n) device elements; its first lag_num_spans(n) elements are left holding the span products of (pt - j);
// host_pinned: >= LAG_SPAN pinned elements; *total_host = prod_j (pt - j).
size_t lag_binv_scratch(size_t n);
cudaError_t launch_node_inverses(const fr_t& pt, size_t n, fr_t* inv, fr_t* scratch, fr_t* host_pinned, fr_t* total_host, int sm_count, cudaStream_t s,
                                 unsigned* launches);
// *out = product of in[0 .. count)   (one block)
cudaError_t launch_fr_product(const fr_t* in, size_t count, fr_t* out, cudaStream_t s);
// *out = prod_{j < m} (pt - j), for m below one span (one thread)
cudaError_t launch_node_product(const fr_t& pt, size_t m, fr_t* out, cudaStream_t s);
// scal[j] = L_{first+j}(tau), j < count, from inv[j] = 1/(tau - first - j), *ntau = prod over ALL n nodes of (tau - k), ifact[k] = 1/k!.  1 launch.
cudaError_t launch_lagrange_scalars(const fr_t* inv, const fr_t* ifact, const fr_t* ntau, size_t n, size_t first, size_t count, fr_t* scal, int sm_count, cudaStream_t s);
// *value = P(z) (barycentric), q[j] = Q(j) for j < n.  partials / ticket: the context's grid-reduction scratch.  2 launches.
cudaError_t launch_bary_open(const fr_t* vals, const fr_t* inv, const fr_t* ifact, size_t n, const fr_t* nz, fr_t* partials, unsigned int* ticket,
                             fr_t* value, fr_t* q, int sm_count, cudaStream_t s);

// sharded forms (a rank holds the nodes first .. first + count - 1 of the n-node domain): raw partial barycentric sum, then the quotient values
cudaError_t launch_bary_partial(const fr_t* vals, const fr_t* inv, const fr_t* ifact, size_t n, size_t first, size_t count, const fr_t* scale,
                                fr_t* partials, unsigned int* ticket, fr_t* partial, int sm_count, cudaStream_t s);
cudaError_t launch_bary_quotient(const fr_t* vals, const fr_t* inv, const fr_t* value, size_t count, fr_t* q, int sm_count, cudaStream_t s);

}  // namespace tsg
