// msm.cu - KZG commitment as a Pippenger multi-scalar multiplication on BN254 G1, plus SRS generation.
//
// Replaces KZGCommitment::commit (src/commitments.rs:162-180: a serial sum of n double-and-add scalar
// multiplications) and the g1_powers loop of setup_params (src/utils.rs:89-96).  The group element
// produced is the same; the algorithm is:
//   1. scalars Montgomery -> canonical, signed base-2^c digits (c <= 16), W = ceil(255 / c) windows
//   2. counting sort of (window, |digit|) -> point index lists (histogram, exclusive scan, scatter)
//   3. bucket accumulation: one thread per work item (a bucket, or a slice of at most MSM_CHUNK entries of
//      an over-full bucket), XYZZ accumulator + affine SRS point = 8M + 2S, sign applied to y
//   4. per-window bucket reduction sum_b b * B_b by blocked running sums, then a tree sum per window
//   5. the W window sums go to the host, which combines them (W * c doublings: microseconds of work)
// Integer-pipe bound: ~ n * W * 10 Fq products; memory traffic is the 64-byte gathers of step 3.
#include <cstdio>
#include <cstdlib>
#include <cooperative_groups.h>
#include "fr_device.cuh"
#include "g1.cuh"
#include <algorithm>
#include "msm.cuh"

namespace tsg {

__device__ __forceinline__ g1_affine ld_affine(const g1_affine* p) {
    g1_affine a;
    a.x = ld256_nc(&p->x); a.y = ld256_nc(&p->y);
    return a;
}
__device__ __forceinline__ g1_xyzz ld_xyzz(const g1_xyzz* p) {
    g1_xyzz a;
    a.X = ld256_cg(&p->X); a.Y = ld256_cg(&p->Y); a.ZZ = ld256_cg(&p->ZZ); a.ZZZ = ld256_cg(&p->ZZZ);
    return a;
}
__device__ __forceinline__ void st_xyzz(g1_xyzz* p, const g1_xyzz& v) {
    st256(&p->X, v.X); st256(&p->Y, v.Y); st256(&p->ZZ, v.ZZ); st256(&p->ZZZ, v.ZZZ);
}

// ---------------------------------------------------------------- 1. digits + histogram
// set_stride = buckets per window when every window owns a bucket set, 0 when all windows of the job share one set (precomputed tables)
__global__ void k_msm_digits(const fr_t* scalars, size_t n, unsigned c, unsigned W, unsigned* dig, unsigned* hist, unsigned set_stride) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const unsigned nb = 1u << (c - 1);
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x; i0 < n; i0 += stride) {   // block-uniform bound: warps stay converged for match_any
        const size_t i = i0 + threadIdx.x;
        const bool live = i < n;
        fr_t s = live ? ld256_nc(scalars + i).from_mont() : fr_t::zero();
        unsigned carry = 0;
        for (unsigned w = 0; w < W; ++w) {
            unsigned bit = w * c, limb = bit >> 5, off = bit & 31;
            unsigned long long two = limb < 8 ? s.l[limb] : 0u;
            if (limb + 1 < 8) two |= (unsigned long long)s.l[limb + 1] << 32;
            unsigned d = (unsigned)((two >> off) & ((1u << c) - 1)) + carry;
            unsigned sign = 0;
            if (d > nb) { d = (1u << c) - d; sign = 1; carry = 1; } else carry = 0;
            if (live) dig[(size_t)w * n + i] = d | (sign << 31);
            // one atomic per distinct bucket in the warp: small scalars put most carries of a window into bucket 1
            const unsigned key = d ? (unsigned)(w * set_stride + d - 1) : 0xffffffffu;
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            if (d && (unsigned)(__ffs(peers) - 1) == (threadIdx.x & 31)) atomicAdd(&hist[key], (unsigned)__popc(peers));
        }
    }
}

// *flag != 0 afterwards iff some scalar (canonical form) does not fit 64 bits
__global__ void k_msm_scalar_probe(const fr_t* scalars, size_t n, unsigned* flag) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned wide = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        fr_t s = ld256_nc(scalars + i);
        if (s.is_zero()) continue;
        s = s.from_mont();
        wide |= s.l[2] | s.l[3] | s.l[4] | s.l[5] | s.l[6] | s.l[7];
    }
    if (__any_sync(0xffffffffu, wide != 0) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}
cudaError_t launch_scalar_probe(const fr_t* scalars, size_t n, unsigned* flag, int sm_count, cudaStream_t s) {
    size_t g = (n + 255) / 256; size_t cap = (size_t)sm_count * 8;
    k_msm_scalar_probe<<<(unsigned)(g < 1 ? 1 : (g < cap ? g : cap)), 256, 0, s>>>(scalars, n, flag);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- 2. exclusive scan of u32 counters (three small kernels)
// tile = SCAN_THREADS * SCAN_ITEMS consecutive counters per block
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ unsigned block_exclusive_scan_u32(unsigned v, unsigned* total) {
    __shared__ unsigned warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += y; }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    unsigned warp_off = warp ? warp_sums[warp - 1] : 0;
    if (total) *total = warp_sums[SCAN_THREADS / 32 - 1];
    return warp_off + x - v;   // exclusive prefix of v within the block
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const unsigned* in, size_t n, unsigned* tile_sums) {
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) if (base + k < n) s += in[base + k];
    unsigned tot;
    block_exclusive_scan_u32(s, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}
// one block: exclusive scan of the tile sums in place (any count), total to *total
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(unsigned* tile_sums, size_t ntiles, unsigned* total) {
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (size_t b = 0; b < ntiles; b += SCAN_THREADS) {
        size_t i = b + threadIdx.x;
        unsigned v = i < ntiles ? tile_sums[i] : 0, tot;
        unsigned ex = block_exclusive_scan_u32(v, &tot);
        unsigned c = carry;
        if (i < ntiles) tile_sums[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const unsigned* in, unsigned* out, size_t n, const unsigned* tile_offsets) {
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    unsigned v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = base + k < n ? in[base + k] : 0; s += v[k]; }
    unsigned acc = tile_offsets[blockIdx.x] + block_exclusive_scan_u32(s, nullptr);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { if (base + k < n) out[base + k] = acc; acc += v[k]; }
}
// out = exclusive scan of in; scratch holds ceil(n / SCAN_TILE) counters
static cudaError_t exclusive_scan_u32(const unsigned* in, unsigned* out, size_t n, unsigned* scratch, unsigned* total, cudaStream_t s) {
    const size_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_tile_sums<<<(unsigned)ntiles, SCAN_THREADS, 0, s>>>(in, n, scratch);
    k_scan_tiles<<<1, SCAN_THREADS, 0, s>>>(scratch, ntiles, total);
    k_scan_apply<<<(unsigned)ntiles, SCAN_THREADS, 0, s>>>(in, out, n, scratch);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- 3. scatter point indices into bucket order
// entry = index into the job's base array (+ sign): i for per-window bucket sets, w * point_stride + i into the precomputed tables
__global__ void k_msm_scatter(const unsigned* dig, size_t n, unsigned c, unsigned W, const unsigned* offsets, unsigned* cursor, unsigned* sorted,
                              unsigned set_stride, size_t point_stride) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t total = (size_t)W * n;
    const unsigned lane = threadIdx.x & 31;
    for (size_t t0 = (size_t)blockIdx.x * blockDim.x; t0 < total; t0 += stride) {   // block-uniform bound: whole warps stay converged
        const size_t t = t0 + threadIdx.x;
        unsigned v = t < total ? dig[t] : 0u;
        unsigned d = v & 0x7fffffffu;
        size_t w = t / n, i = t - w * n;
        const unsigned b = d ? (unsigned)(w * set_stride + d - 1) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, b);
        const unsigned leader = __ffs(peers) - 1;
        unsigned base = 0;
        if (d && lane == leader) base = atomicAdd(&cursor[b], (unsigned)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (d) sorted[offsets[b] + base + __popc(peers & ((1u << lane) - 1))] = (unsigned)(w * point_stride + i) | (v & 0x80000000u);
    }
}

// ---------------------------------------------------------------- work items: slices of at most MSM_CHUNK entries
// max_chunks[0] = largest chunk count of a bucket, max_chunks[1] = number of buckets with more than MSM_SERIAL_MERGE chunks, listed in heavy[] (any order)
__global__ void k_msm_item_counts(const unsigned* hist, size_t nbuckets, unsigned chunk, unsigned* items, unsigned* max_chunks, unsigned* heavy) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned mx = 0;
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nbuckets; b += stride) {
        unsigned k = (hist[b] + chunk - 1) / chunk;
        items[b] = k;
        mx = k > mx ? k : mx;
        if (k > MSM_SERIAL_MERGE) heavy[atomicAdd(max_chunks + 1, 1u)] = (unsigned)b;
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx > 1) atomicMax(max_chunks, mx);
}
__global__ void __launch_bounds__(256) k_msm_item_fill(const unsigned* items, const unsigned* item_off, const unsigned* hist, size_t nbuckets, unsigned chunk,
                                                       unsigned* item_bucket, unsigned* len_hist) {
    // len_hist[MSM_CHUNK - len] counts work items by length (longest first), privatised in shared memory
    __shared__ unsigned sh[MSM_CHUNK + 1];
    for (unsigned i = threadIdx.x; i <= MSM_CHUNK; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const unsigned lane = threadIdx.x & 31;
    for (size_t b0 = (size_t)blockIdx.x * blockDim.x; b0 < nbuckets; b0 += stride) {
        const size_t b = b0 + threadIdx.x;
        unsigned k = 0, o = 0, cnt = 0;
        if (b < nbuckets) { k = items[b]; o = item_off[b]; cnt = hist[b]; }
        if (k <= 8) {
            for (unsigned j = 0; j < k; ++j) {
                item_bucket[o + j] = (unsigned)b;
                unsigned len = cnt - j * chunk; if (len > chunk) len = chunk;
                atomicAdd(&sh[MSM_CHUNK - len], 1u);
            }
        }
        // buckets split into many chunks (equal scalars, carries of small scalars): the whole warp writes their item records
        unsigned heavy = __ballot_sync(0xffffffffu, k > 8);
        while (heavy) {
            const int src = __ffs(heavy) - 1; heavy &= heavy - 1;
            const unsigned kb = __shfl_sync(0xffffffffu, k, src), ob = __shfl_sync(0xffffffffu, o, src), cb = __shfl_sync(0xffffffffu, cnt, src);
            const unsigned bb = (unsigned)(b0 + (threadIdx.x & ~31u) + src);
            for (unsigned j = lane; j < kb; j += 32) {
                item_bucket[ob + j] = bb;
                unsigned len = cb - j * chunk; if (len > chunk) len = chunk;
                atomicAdd(&sh[MSM_CHUNK - len], 1u);
            }
        }
    }
    __syncthreads();
    for (unsigned i = threadIdx.x; i <= MSM_CHUNK; i += blockDim.x) if (sh[i]) atomicAdd(&len_hist[i], sh[i]);
}
// exclusive scan of the MSM_CHUNK + 1 length bins (one block)
__global__ void __launch_bounds__(SCAN_THREADS) k_msm_len_scan(const unsigned* len_hist, unsigned* len_off) {
    unsigned v0 = 2 * threadIdx.x <= MSM_CHUNK ? len_hist[2 * threadIdx.x] : 0;
    unsigned v1 = 2 * threadIdx.x + 1 <= MSM_CHUNK ? len_hist[2 * threadIdx.x + 1] : 0;
    unsigned ex = block_exclusive_scan_u32(v0 + v1, nullptr);
    if (2 * threadIdx.x <= MSM_CHUNK) len_off[2 * threadIdx.x] = ex;
    if (2 * threadIdx.x + 1 <= MSM_CHUNK) len_off[2 * threadIdx.x + 1] = ex + v0;
}
// order[] = work items sorted by decreasing length, so that the 32 lanes of a warp of k_msm_accumulate run the
// same number of additions (bucket loads are Poisson distributed: unsorted, a warp waits for its fullest bucket)
constexpr int ORDER_TILE = 4;   // items per thread per tile
__global__ void __launch_bounds__(256) k_msm_order(const unsigned* hist, const unsigned* item_off, const unsigned* item_bucket, const unsigned* n_items, unsigned chunk,
                                                   const unsigned* len_off, unsigned* len_cursor, unsigned* order) {
    __shared__ unsigned sh_cnt[MSM_CHUNK + 1], sh_base[MSM_CHUNK + 1];
    const unsigned M = *n_items;
    const unsigned tile = blockDim.x * ORDER_TILE;
    for (unsigned t0 = blockIdx.x * tile; t0 < M; t0 += gridDim.x * tile) {
        for (unsigned i = threadIdx.x; i <= MSM_CHUNK; i += blockDim.x) sh_cnt[i] = 0;
        __syncthreads();
        unsigned key[ORDER_TILE], rank[ORDER_TILE];
#pragma unroll
        for (int q = 0; q < ORDER_TILE; ++q) {
            unsigned it = t0 + q * blockDim.x + threadIdx.x;
            key[q] = 0xffffffffu;
            if (it < M) {
                unsigned b = item_bucket[it];
                unsigned len = hist[b] - (it - item_off[b]) * chunk; if (len > chunk) len = chunk;
                key[q] = MSM_CHUNK - len;
                rank[q] = atomicAdd(&sh_cnt[key[q]], 1u);
            }
        }
        __syncthreads();
        for (unsigned i = threadIdx.x; i <= MSM_CHUNK; i += blockDim.x) if (sh_cnt[i]) sh_base[i] = len_off[i] + atomicAdd(&len_cursor[i], sh_cnt[i]);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < ORDER_TILE; ++q) if (key[q] != 0xffffffffu) order[sh_base[key[q]] + rank[q]] = t0 + q * blockDim.x + threadIdx.x;
        __syncthreads();
    }
}

// ---------------------------------------------------------------- 3b. bucket accumulation (the hot kernel)
// Register allocation: without a min-blocks bound ptxas settles on 122 registers (4 blocks of 128 threads per SM: up to 128 registers fit four), which measured best:
// 96 registers / 5 blocks (small spills) 4.86 ms, 80 / 6 blocks 5.00 ms, 130 / 3 blocks 5.00 ms against 4.81 ms per 2^20-op proof
__global__ void __launch_bounds__(MSM_ACC_THREADS) k_msm_accumulate(const MsmBases jobs, unsigned buckets_per_job, const unsigned* sorted, const unsigned* hist, const unsigned* offsets,
                                                                  const unsigned* item_off, const unsigned* item_bucket, const unsigned* n_items, unsigned chunk,
                                                                  const unsigned* order, g1_xyzz* partial) {
    const unsigned M = *n_items;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += stride) {
        const unsigned it = order[j];
        unsigned b = item_bucket[it];
        unsigned k = it - item_off[b];
        unsigned cnt = hist[b], base = offsets[b];
        unsigned lo = k * chunk, hi = lo + chunk < cnt ? lo + chunk : cnt;
        const g1_affine* bases = jobs.p[b / buckets_per_job];
        g1_xyzz acc = g1_xyzz::identity();
        for (unsigned p = lo; p < hi; ++p) {
            unsigned v = sorted[base + p];
            g1_affine pt = ld_affine(bases + (v & 0x7fffffffu));
            acc = acc.add_affine(pt, (v >> 31) != 0);
        }
        st_xyzz(partial + it, acc);
    }
}

// ---------------------------------------------------------------- 3c'. buckets split into a few chunks: one thread per bucket adds its chunks into the first one
// (the common case: Poisson tails, the piled-up top window, the short chunks of small passes - a chain of 1-2 additions on every split bucket at once;
// the cooperative tree below, a full pass + grid barrier per level, stays for buckets with many chunks: equal scalars, carries of small scalars)
__global__ void __launch_bounds__(128) k_msm_merge_serial(g1_xyzz* partial, const unsigned* items, const unsigned* item_off, size_t nbuckets, const unsigned* n_items) {
    const unsigned max_chunks = n_items[2];
    if (max_chunks <= 1 || max_chunks > MSM_BLOCK_MERGE) return;   // beyond MSM_BLOCK_MERGE the cooperative tree merges every bucket
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nbuckets; b += stride) {
        const unsigned k = items[b];
        if (k < 2 || k > MSM_SERIAL_MERGE) continue;               // heavier buckets: k_msm_merge_heavy
        g1_xyzz* pb = partial + item_off[b];
        g1_xyzz acc = ld_xyzz(pb);
        for (unsigned q = 1; q < k; ++q) acc = acc.add(ld_xyzz(pb + q));
        st_xyzz(pb, acc);
    }
}

// ---------------------------------------------------------------- 3c. buckets split into several chunks: pairwise tree over their partial sums
// round with stride s: chunk q of a bucket (q a multiple of 2s) absorbs chunk q + s; after ceil(log2 k) rounds chunk 0 holds the bucket sum.
// One cooperative launch runs every round (grid.sync between rounds) and returns at once when no bucket was split.
__global__ void __launch_bounds__(128) k_msm_merge_chunks(g1_xyzz* partial, const unsigned* items, const unsigned* item_off, const unsigned* item_bucket,
                                                          const unsigned* n_items) {
    const unsigned max_chunks = n_items[2];    // largest chunk count of any bucket (0 when none exceeds one chunk)
    if (max_chunks <= MSM_BLOCK_MERGE) return;   // k_msm_merge_serial / k_msm_merge_heavy have merged every split bucket
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const unsigned M = n_items[0];
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (unsigned s = 1; s < max_chunks; s <<= 1) {
        for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < M; it += stride) {
            const unsigned b = item_bucket[it];
            const unsigned k = items[b];
            if (k <= s) continue;
            const unsigned q = (unsigned)it - item_off[b];
            if ((q & (2 * s - 1)) == 0 && q + s < k) st_xyzz(partial + it, ld_xyzz(partial + it).add(ld_xyzz(partial + it + s)));
        }
        grid.sync();
    }
}

// ---------------------------------------------------------------- lane-cooperative XYZZ addition (the latency-bound tree levels of the reduction)
// A lone thread needs ~14 dependent field products (~3500 instructions, 10-20 us when its warp has the scheduler to itself) for one XYZZ + XYZZ
// addition; tree levels run few of them, so their cost is that latency.  Here the four lanes of a QUAD (lanes 4q .. 4q + 3 of a warp) share one
// addition of add-2008-s: four product steps with one product per lane, operands exchanged through QUAD_SLOTS field elements of shared memory:
//   step 1   U1 = X1 ZZ2 | U2 = X2 ZZ1 | S1 = Y1 ZZZ2 | S2 = Y2 ZZZ1          (then every lane: P = U2 - U1, R = S2 - S1)
//   step 2   PP = P^2    | ZZ1 ZZ2     | R^2          | ZZZ1 ZZZ2
//   step 3   Q = U1 PP   | PPP = P PP  | ZZ3 = ZZ1 ZZ2 PP | (PPP)
//   step 4   X3 = R^2 - PPP - 2Q, Y3 = R (Q - X3) - S1 PPP | ZZZ3 = ZZZ1 ZZZ2 PPP      (one fused a b - c d on every lane: no divergence)
// i.e. ~4.5 products deep instead of ~13.5.  Same bits as g1_xyzz::add.  All 32 lanes of a warp must call it (quads with active = false only keep the
// warp converged: their pointers must still be readable); dst may alias A or B; the caller orders successive tree levels (__syncthreads / __syncwarp).
constexpr int QUAD_SLOTS = 12;
__device__ __noinline__ void xyzz_add_alone(g1_xyzz* dst, const g1_xyzz* A, const g1_xyzz* B) {
    const g1_xyzz a = *A, b = *B;
    *dst = a.add(b);
}
__device__ __forceinline__ fq_t fq_sel(bool c, const fq_t& a, const fq_t& b) {
    fq_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = c ? a.l[i] : b.l[i];
    return r;
}
__device__ __forceinline__ void quad_add(g1_xyzz* dst, const g1_xyzz* A, const g1_xyzz* B, fq_t* scr, bool active) {
    const unsigned r = threadIdx.x & 3;
    const fq_t* fa = reinterpret_cast<const fq_t*>(A);   // fields in order X, Y, ZZ, ZZZ
    const fq_t* fb = reinterpret_cast<const fq_t*>(B);
    fq_t* fd = reinterpret_cast<fq_t*>(dst);
    const bool idA = fa[2].is_zero(), idB = fb[2].is_zero();
    {   // step 1
        const fq_t* s0 = (r & 1) ? fb : fa; const fq_t* s1 = (r & 1) ? fa : fb;
        const fq_t a = s0[r >> 1], b = s1[2 + (r >> 1)];
        scr[r] = a * b;
    }
    __syncwarp();
    const fq_t P = scr[1] - scr[0], R = scr[3] - scr[2];
    const bool same_x = P.is_zero();
    {   // step 2
        const fq_t ma = fa[2 + (r >> 1)], mb = fb[2 + (r >> 1)];
        const fq_t a = fq_sel(r == 0, P, fq_sel(r == 2, R, ma)), b = fq_sel(r == 0, P, fq_sel(r == 2, R, mb));
        scr[4 + r] = a * b;   // 4: PP, 5: ZZ1 ZZ2, 6: R^2, 7: ZZZ1 ZZZ2
    }
    __syncwarp();
    {   // step 3
        const fq_t PP = scr[4];
        const fq_t a = fq_sel(r == 0, scr[0], fq_sel(r == 2, scr[5], P));
        scr[8 + r] = a * PP;   // 8: Q, 9: PPP, 10: ZZ3, 11: PPP again
    }
    __syncwarp();
    {   // step 4
        const fq_t Q = scr[8], PPP = scr[9];
        const fq_t X3 = scr[6] - PPP - Q.dbl();
        const fq_t a = fq_sel(r == 0, R, scr[7]), b = fq_sel(r == 0, Q - X3, PPP), c = fq_sel(r == 0, scr[2], fq_t::zero());
        const fq_t res = fq_t::mul_sub(a, b, c, PPP);
        if (active) {
            if (idB) { if (dst != A) fd[r] = fa[r]; }
            else if (idA) fd[r] = fb[r];
            else if (same_x) { if (r == 0) xyzz_add_alone(dst, A, B); }   // doubling / inverse points: the lone-thread formulas
            else if (r == 0) { fd[0] = X3; fd[1] = res; }
            else if (r == 1) fd[3] = res;
            else if (r == 2) fd[2] = scr[10];
        }
    }
    __syncwarp();
}
// tree sum of sh[0 .. n) (n a power of two) into sh[0] by the quads of a block; scr_all: QUAD_SLOTS elements per quad.  Block-wide: every thread calls.
__device__ __forceinline__ void quad_tree_sum(g1_xyzz* sh, unsigned n, fq_t* scr_all) {
    const unsigned q = threadIdx.x >> 2, quads = blockDim.x >> 2;
    fq_t* scr = scr_all + (size_t)q * QUAD_SLOTS;
    for (unsigned s = n >> 1; s > 0; s >>= 1) {
        for (unsigned base = 0; base < s; base += quads) {
            if (base + ((threadIdx.x >> 5) << 3) < s) {   // warp-uniform: some quad of this warp has a pair
                const unsigned pair = base + q;
                const bool act = pair < s;
                const unsigned i = act ? pair : 0;
                quad_add(sh + i, sh + i, sh + i + s, scr, act);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- 3c''. buckets split into MANY chunks (but at most MSM_BLOCK_MERGE): one BLOCK per listed bucket
// The piled-up top window of a large MSM (2^24 points, c = 22: 2^24 entries on 2^11 buckets = 128 chunks each) and short scalars with a narrow top window land here.
// Every thread adds a strided share of the bucket's partial sums, the 128 results meet in a lane-quad tree in shared memory.  The cooperative tree, which scans ALL work
// items once per level (7 levels x 5 M items at 2^24 points: 1.9 ms), is left for the degenerate passes (equal scalars: one bucket with thousands of chunks).
constexpr int MSM_HEAVY_THREADS = 128;
__global__ void __launch_bounds__(MSM_HEAVY_THREADS) k_msm_merge_heavy(g1_xyzz* partial, const unsigned* items, const unsigned* item_off, const unsigned* heavy, const unsigned* n_items) {
    const unsigned max_chunks = n_items[2], n_heavy = n_items[3];
    if (max_chunks <= MSM_SERIAL_MERGE || max_chunks > MSM_BLOCK_MERGE) return;
    __shared__ g1_xyzz sh[MSM_HEAVY_THREADS];
    __shared__ fq_t scr[(MSM_HEAVY_THREADS / 4) * QUAD_SLOTS];
    for (unsigned h = blockIdx.x; h < n_heavy; h += gridDim.x) {
        const unsigned b = heavy[h], k = items[b];
        g1_xyzz* pb = partial + item_off[b];
        g1_xyzz acc = g1_xyzz::identity();
        for (unsigned i = threadIdx.x; i < k; i += MSM_HEAVY_THREADS) acc = acc.add(ld_xyzz(pb + i));
        sh[threadIdx.x] = acc;
        __syncthreads();
        unsigned m = 1; while (m < k && m < (unsigned)MSM_HEAVY_THREADS) m <<= 1;   // block-uniform: the tree spans the slots that hold a sum
        quad_tree_sum(sh, m, scr);   // ends on a block barrier
        if (threadIdx.x == 0) st_xyzz(pb, sh[0]);
        __syncthreads();
    }
}

// ---------------------------------------------------------------- 4. window reduction  S_w = sum_b (b + 1) * B_{w,b}
// Three short launches with a shallow dependency chain (a lone XYZZ addition has ~6 us latency, so depth is what costs):
//   a. span sums: thread per span of S consecutive buckets: R = sum B, L = sum_j (j + 1) B_{lo + j}   (2 S additions deep)
//      S_w = sum_spans L + S * sum_sp sp * R_sp
//   b. the index-weighted sum by bits of sp: sum_sp sp R_sp = sum_k 2^k (sum over sp with bit k set of R_sp): one block per
//      (window, bit) tree-sums its half of the R's and doubles the result k + log2 S times; one more block per window sums the L's
//   c. one warp-sized block per window adds the (bits + 1) block results and writes the window sum as a Jacobian point
__global__ void __launch_bounds__(128) k_msm_span_sums(const g1_xyzz* partial, const unsigned* items, const unsigned* item_off, unsigned nb, unsigned S,
                                                       size_t total_spans, g1_xyzz* R, g1_xyzz* L) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const unsigned T = nb / S;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total_spans; t += stride) {
        const size_t w = t / T; const unsigned lo = (unsigned)(t - w * T) * S;
        g1_xyzz running = g1_xyzz::identity(), acc = g1_xyzz::identity();
        for (unsigned j = S; j-- > 0;) {
            const size_t b = w * nb + lo + j;
            if (items[b]) running = running.add(ld_xyzz(partial + item_off[b]));
            acc = acc.add(running);
        }
        st_xyzz(R + t, running); st_xyzz(L + t, acc);
    }
}
__global__ void __launch_bounds__(MSM_SUM_THREADS) k_msm_bit_sums(const g1_xyzz* R, const g1_xyzz* L, unsigned T, unsigned nbits, unsigned k0, g1_xyzz* parts, int quad) {
    // block (k, w, p): slice p of gridDim.z of the elements that bit k selects (k == nbits: all L's) of bucket set w
    extern __shared__ __align__(16) unsigned char bit_sums_smem[];
    g1_xyzz* sh = reinterpret_cast<g1_xyzz*>(bit_sums_smem);                       // MSM_SUM_THREADS points, then the quads' exchange slots
    fq_t* scr = reinterpret_cast<fq_t*>(sh + MSM_SUM_THREADS);
    // slots k < nbits: the R's whose span index has bit k set; slots nbits and nbits + 1: the two halves of the L's (so that every
    // block sums the same number of elements, T / 2 / P)
    const unsigned k = blockIdx.x + k0, w = blockIdx.y, p = blockIdx.z, P = gridDim.z;   // k0: first slot of this launch
    const unsigned U = T > 1 ? T / 2 : (k == nbits ? 1 : 0);
    const unsigned u0 = (unsigned)((unsigned long long)U * p / P), u1 = (unsigned)((unsigned long long)U * (p + 1) / P);
    g1_xyzz acc = g1_xyzz::identity();
    if (k >= nbits) {
        const unsigned base = k == nbits ? 0 : U;
        for (unsigned i = u0 + threadIdx.x; i < u1; i += blockDim.x) acc = acc.add(ld_xyzz(L + (size_t)w * T + base + i));
    } else {
        for (unsigned u = u0 + threadIdx.x; u < u1; u += blockDim.x) {
            const unsigned t = ((u >> k) << (k + 1)) | (1u << k) | (u & ((1u << k) - 1));
            acc = acc.add(ld_xyzz(R + (size_t)w * T + t));
        }
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    if (quad) {
        quad_tree_sum(sh, blockDim.x, scr);
    } else {
        for (unsigned s = blockDim.x / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) sh[threadIdx.x] = sh[threadIdx.x].add(sh[threadIdx.x + s]);
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) st_xyzz(parts + ((size_t)w * (nbits + 2) + k) * P + p, sh[0]);
}
// block (k, w): adds the P <= 32 slice sums of (set w, bit k) and writes the sum as a Jacobian point.  The weights 2^(k + log2 S) and the
// final additions are a Horner evaluation over span_bits + 1 points per set: ~40 group operations, done by the host in microseconds
// (on the device they would be a chain of ~20 dependent doublings, ~0.1 ms of pure latency)
__global__ void __launch_bounds__(MSM_FIN_THREADS) k_msm_bit_finish(const g1_xyzz* parts, unsigned P, unsigned nbits, unsigned k0, g1_jac* out, int quad) {
    __shared__ g1_xyzz sh[32];
    __shared__ fq_t scr[(MSM_FIN_THREADS / 4) * QUAD_SLOTS];
    const unsigned k = blockIdx.x + k0, w = blockIdx.y;
    const size_t slot = (size_t)w * (nbits + 2) + k;
    if (threadIdx.x < 32) sh[threadIdx.x] = threadIdx.x < P ? ld_xyzz(parts + slot * P + threadIdx.x) : g1_xyzz::identity();
    __syncthreads();
    if (quad) {
        unsigned n = 1; while (n < P) n <<= 1;
        quad_tree_sum(sh, n, scr);
        // (X ZZ^2, Y ZZZ^2, ZZZ): lanes 0 and 1 take one coordinate each (same instructions, other operands)
        if (threadIdx.x < 2) {
            const fq_t* f = reinterpret_cast<const fq_t*>(&sh[0]);
            const bool id = f[2].is_zero();
            const fq_t z = f[2 + threadIdx.x], c = f[threadIdx.x];
            const fq_t v = id ? fq_t::one() : c * z.sqr();
            st256(threadIdx.x ? &out[slot].y : &out[slot].x, v);
            if (threadIdx.x == 0) st256(&out[slot].z, id ? fq_t::zero() : f[3]);
        }
        return;
    }
    if (threadIdx.x >= 32) return;
    for (unsigned s = 16; s > 0; s >>= 1) {
        if (threadIdx.x < s && threadIdx.x + s < P) sh[threadIdx.x] = sh[threadIdx.x].add(sh[threadIdx.x + s]);
        __syncwarp();
    }
    if (threadIdx.x == 0) {
        g1_jac j = sh[0].to_jacobian();
        st256(&out[slot].x, j.x); st256(&out[slot].y, j.y); st256(&out[slot].z, j.z);
    }
}

// ---------------------------------------------------------------- SRS generation: out[i] = tau^i * G
// powers of tau: thread handles MSM_POW_SPAN consecutive exponents
__global__ void k_tau_powers(const fr_t tau, size_t first, size_t n, fr_t* out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t nchunks = (n + MSM_POW_SPAN - 1) / MSM_POW_SPAN;
    for (size_t ch = (size_t)blockIdx.x * blockDim.x + threadIdx.x; ch < nchunks; ch += stride) {
        size_t b = ch * MSM_POW_SPAN, e = b + MSM_POW_SPAN < n ? b + MSM_POW_SPAN : n;
        fr_t cur = tau.pow_u64(first + b);
        for (size_t i = b; i < e; ++i) { st256(out + i, cur); cur = cur * tau; }
    }
}
// fixed-base multiplication with a byte-window table: table[w * 255 + d - 1] = d * 256^w * G (affine)
__global__ void __launch_bounds__(128) k_fixed_base_mul(const fr_t* scalars, size_t n, const g1_affine* table, g1_xyzz* out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        fr_t s = ld256_nc(scalars + i).from_mont();
        g1_xyzz acc = g1_xyzz::identity();
#pragma unroll 1
        for (int w = 0; w < 32; ++w) {
            unsigned d = (s.l[w >> 2] >> (8 * (w & 3))) & 0xffu;
            if (d) acc = acc.add_affine(ld_affine(table + w * 255 + d - 1));
        }
        st_xyzz(out + i, acc);
    }
}

// XYZZ -> affine with Montgomery batch inversion over MSM_INV_SPAN points per thread
__global__ void __launch_bounds__(128) k_batch_to_affine(const g1_xyzz* in, size_t n, g1_affine* out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t nchunks = (n + MSM_INV_SPAN - 1) / MSM_INV_SPAN;
    for (size_t ch = (size_t)blockIdx.x * blockDim.x + threadIdx.x; ch < nchunks; ch += stride) {
        size_t b = ch * MSM_INV_SPAN, e = b + MSM_INV_SPAN < n ? b + MSM_INV_SPAN : n;
        // forward: out[i].x temporarily holds the prefix product of the ZZZ's before i
        fq_t acc = fq_t::one();
        for (size_t i = b; i < e; ++i) {
            st256(&out[i].x, acc);
            fq_t z = ld256_nc(&in[i].ZZZ);
            if (!z.is_zero()) acc = acc * z;
        }
        fq_t inv = acc.inverse();
        for (size_t i = e; i-- > b;) {
            g1_xyzz p; p.X = ld256_nc(&in[i].X); p.Y = ld256_nc(&in[i].Y); p.ZZ = ld256_nc(&in[i].ZZ); p.ZZZ = ld256_nc(&in[i].ZZZ);
            if (p.ZZZ.is_zero()) { st256(&out[i].x, fq_t::zero()); st256(&out[i].y, fq_t::zero()); continue; }
            fq_t pref = ld256(&out[i].x);
            fq_t t = inv * pref;          // 1 / ZZZ_i
            inv = inv * p.ZZZ;
            fq_t u = p.ZZ * t;            // 1 / Z_i
            st256(&out[i].x, p.X * u.sqr());
            st256(&out[i].y, p.Y * t);
        }
    }
}

// Jacobian {x,y,z} (reference G1Projective) -> XYZZ, for SRS upload
__global__ void k_jac_to_xyzz(const g1_jac* in, size_t n, g1_xyzz* out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        g1_jac j; j.x = ld256_nc(&in[i].x); j.y = ld256_nc(&in[i].y); j.z = ld256_nc(&in[i].z);
        st_xyzz(out + i, g1_xyzz::from_jacobian(j));
    }
}
// affine -> Jacobian with z = 1 (identity -> z = 0), for returning g1_powers to the host
__global__ void k_affine_to_jac(const g1_affine* in, size_t n, g1_jac* out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        g1_affine a = ld_affine(in + i);
        bool id = a.is_identity();
        st256(&out[i].x, id ? fq_t::one() : a.x); st256(&out[i].y, id ? fq_t::one() : a.y); st256(&out[i].z, id ? fq_t::zero() : fq_t::one());
    }
}

// ---------------------------------------------------------------- launchers
static int g_msm_quad_tree = 1;   // tree levels of the window reduction by lane quads (tuning "msm_quad_tree"; 0 = one thread per addition)
void msm_set_quad_tree(int on) { g_msm_quad_tree = on ? 1 : 0; }
static inline int gridfor(size_t work, int threads, size_t cap) {
    size_t g = (work + threads - 1) / threads;
    if (g < 1) g = 1;
    return (int)(g < cap ? g : cap);
}

unsigned msm_window_bits(size_t n) {
    unsigned lg = 0; while (((size_t)1 << (lg + 1)) <= n) ++lg;
    int c = (int)lg - 3;
    if (c < 4) c = 4;
    if (c > 16) c = 16;
    return (unsigned)c;
}

// window width of the precomputed-table mode for a base array of n points
unsigned msm_table_window_bits(size_t n) {
    if (const char* env = getenv("TSGPU_TABLE_WINDOW_BITS")) { int v = atoi(env); if (v >= 4 && v <= 24) return (unsigned)v; }   // experiments
    unsigned lg = 0; while (((size_t)1 << (lg + 1)) <= n) ++lg;
    if (lg < 8) lg = 8;
    if (lg >= 24) return 22;   // 12 windows; 2^21 buckets still hold ~100 entries each
    if (lg >= 19) return 20;   // 13 windows (21 would not save one).  2^19 points: measured 5.58 ms per sharded-proof shape against 6.02 ms with c = 19 (14 windows whose top one
                               // covers 7 bits of the scalar and piles 2^19 entries onto ~100 buckets)
    if (lg == 18) return 17;   // c = 18 needs the same 15 windows as 17 and its top window covers 2 bits (all 2^18 entries in 3 buckets: a 12-level chunk merge): 3.53 against 3.77 ms
    return lg;
}

size_t msm_scratch_bytes(size_t nmax, int K, unsigned c, bool shared, MsmLayout* L, unsigned windows) {
    const unsigned W = windows ? windows : (255 + c - 1) / c;   // windows: only the low digit positions are scanned (scalars known to be short)
    const unsigned sets = (unsigned)K * (shared ? 1u : W);            // bucket sets = windows seen by the reduction
    const size_t nb = (size_t)1 << (c - 1), nbuckets = sets * nb;
    const size_t entries = (size_t)K * W * nmax;
    // Work-item length.  Many buckets (>= MSM_ITEMS_TARGET): one item per bucket is parallelism enough, long chunks keep the merges rare.  Few buckets (the per-rank shapes of a
    // sharded proof: 2^16 buckets per set): shorter chunks, so that the accumulation has ~3 waves of items to balance (2^17 ops per rank: 0.89 -> 0.68 ms) - as long as a bucket
    // still splits into <= ~3 chunks, which k_msm_merge_serial adds with one thread per bucket
    unsigned chunk = MSM_CHUNK;
    if (const char* env = getenv("TSGPU_MSM_CHUNK")) { int v = atoi(env); if (v >= 1 && v <= (int)MSM_CHUNK) chunk = (unsigned)v; }   // experiments
    else while (chunk > 16 && std::max(nbuckets, entries / chunk) + nbuckets / 2 < MSM_ITEMS_TARGET && entries / nbuckets < (size_t)chunk) chunk >>= 1;
    const size_t max_items = nbuckets + entries / chunk + 1;
    // span of the window reduction: long spans when there are many buckets (throughput-bound), short ones when the dependency
    // chain of 2 x span additions would dominate (aim at >= 32768 span threads)
    unsigned span = MSM_RED_SPAN;
    size_t min_spans = MSM_RED_MIN_SPANS;
    if (const char* env = getenv("TSGPU_RED_SPAN")) { int v = atoi(env); if (v >= 2 && v <= 64 && !(v & (v - 1))) span = (unsigned)v; }          // experiments
    if (const char* env = getenv("TSGPU_RED_MIN_SPANS")) { long v = atol(env); if (v >= 1) min_spans = (size_t)v; }
    while (span > 2 && nbuckets / span < min_spans) span >>= 1;
    if (span > nb) span = (unsigned)nb;
    unsigned nbits = 0; while (((size_t)span << nbits) < nb) ++nbits;   // spans per bucket set = 2^nbits
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    L->c = c; L->W = W; L->K = (unsigned)K; L->sets = sets; L->shared = shared; L->nmax = nmax; L->chunk = chunk;
    L->nbuckets = nbuckets; L->max_items = max_items; L->blocks_per_window = (unsigned)(nb / span); L->span = span; L->span_bits = nbits;
    L->dig = take(entries * 4);
    L->sorted = take(entries * 4);
    L->hist = take(nbuckets * 4);                  // hist .. len_hist are contiguous: msm_run clears them with one memset
    L->cursor = take(nbuckets * 4);
    L->n_items = take(256);
    L->len_hist = take(3 * (MSM_CHUNK + 1) * 4);   // length histogram, offsets, cursors
    L->zero_bytes = off - L->hist;
    L->offsets = take(nbuckets * 4);
    L->heavy = take(nbuckets * 4);                 // buckets with more than MSM_SERIAL_MERGE chunks (k_msm_item_counts)
    L->items = take(nbuckets * 4);
    L->item_off = take(nbuckets * 4);
    L->item_bucket = take(max_items * 4);
    L->order = take(max_items * 4);
    L->scan_tmp = take((nbuckets / 1024 + 64) * 4);
    L->partial = take(max_items * sizeof(g1_xyzz));
    L->blockres = take((size_t)2 * sets * L->blocks_per_window * sizeof(g1_xyzz));   // span sums R, then L
    L->bits = take(((size_t)sets * (nbits + 2) * 32 + (size_t)sets * (((size_t)1 << (nbits / 2)) + ((size_t)1 << (nbits - nbits / 2)))) * sizeof(g1_xyzz));   // up to 32 slice sums per (set, slot), then the row / column sums of the two-level form
    L->window_out = take((size_t)sets * (nbits + 2) * sizeof(g1_jac));   // per set: span_bits index-bit sums, then the sums of the two halves of the L's
    return off;
}

// K independent MSMs in one pass (same window width): their bucket sets are laid side by side, so every phase after the
// per-job digit extraction / scatter runs once over the union and the latency-bound tails (chunk chains, tree merge, window
// reduction) overlap across jobs.  Results at scratch + L.window_out, per bucket set (job-major) span_bits + 2 Jacobian points:
// the sums selected by each bit of the span index, then the two halves of the sum of the span-local weighted sums; msm_combine_set() finishes on the host.
cudaError_t msm_run(const MsmJob* jobs, int K, const MsmLayout& L, unsigned char* scratch, int sm_count,
                    cudaStream_t s, unsigned* launches, cudaEvent_t* ev) {
    unsigned* dig = (unsigned*)(scratch + L.dig); unsigned* sorted = (unsigned*)(scratch + L.sorted);
    unsigned* hist = (unsigned*)(scratch + L.hist); unsigned* offsets = (unsigned*)(scratch + L.offsets);
    unsigned* cursor = (unsigned*)(scratch + L.cursor); unsigned* items = (unsigned*)(scratch + L.items);
    unsigned* item_off = (unsigned*)(scratch + L.item_off); unsigned* item_bucket = (unsigned*)(scratch + L.item_bucket);
    unsigned* n_items = (unsigned*)(scratch + L.n_items);
    g1_xyzz* partial = (g1_xyzz*)(scratch + L.partial); g1_xyzz* blockres = (g1_xyzz*)(scratch + L.blockres);
    g1_jac* wout = (g1_jac*)(scratch + L.window_out);
    cudaError_t e;
    if (ev) cudaEventRecord(ev[0], s);
    // bucket histogram, scatter cursors, counters and the length histogram lie side by side: one clear
    if ((e = cudaMemsetAsync(hist, 0, L.zero_bytes, s))) return e;
    unsigned* order = (unsigned*)(scratch + L.order);
    unsigned* len_hist = (unsigned*)(scratch + L.len_hist); unsigned* len_off = len_hist + (MSM_CHUNK + 1); unsigned* len_cursor = len_off + (MSM_CHUNK + 1);
    const size_t cap = (size_t)sm_count * 8;
    const unsigned nb = 1u << (L.c - 1);
    const unsigned set_stride = L.shared ? 0u : nb;
    const size_t buckets_per_job = (size_t)(L.sets / L.K) * nb;
    MsmBases bases;
    for (int k = 0; k < MSM_MAX_BATCH; ++k) bases.p[k] = k < K ? jobs[k].bases : nullptr;
    for (int k = 0; k < K; ++k)
        k_msm_digits<<<gridfor(jobs[k].n, 256, cap), 256, 0, s>>>(jobs[k].scalars, jobs[k].n, L.c, L.W, dig + (size_t)k * L.W * L.nmax,
                                                                   hist + k * buckets_per_job, set_stride);
    exclusive_scan_u32(hist, offsets, L.nbuckets, (unsigned*)(scratch + L.scan_tmp), n_items + 1, s);   // n_items[1] = bucket entries (non-zero digits)
    for (int k = 0; k < K; ++k)
        k_msm_scatter<<<gridfor((size_t)L.W * jobs[k].n, 256, cap), 256, 0, s>>>(dig + (size_t)k * L.W * L.nmax, jobs[k].n, L.c, L.W, offsets + k * buckets_per_job,
                                                                                 cursor + k * buckets_per_job, sorted, set_stride, L.shared ? jobs[k].stride : 0);
    k_msm_item_counts<<<gridfor(L.nbuckets, 256, cap), 256, 0, s>>>(hist, L.nbuckets, L.chunk, items, n_items + 2, (unsigned*)(scratch + L.heavy));
    exclusive_scan_u32(items, item_off, L.nbuckets, (unsigned*)(scratch + L.scan_tmp), n_items, s);
    k_msm_item_fill<<<gridfor(L.nbuckets, 256, cap), 256, 0, s>>>(items, item_off, hist, L.nbuckets, L.chunk, item_bucket, len_hist);
    k_msm_len_scan<<<1, SCAN_THREADS, 0, s>>>(len_hist, len_off);
    k_msm_order<<<gridfor(L.max_items, 256 * ORDER_TILE, cap), 256, 0, s>>>(hist, item_off, item_bucket, n_items, L.chunk, len_off, len_cursor, order);
    if (ev) cudaEventRecord(ev[1], s);
    {
        // 16 blocks per SM, scheduled dynamically: measured faster than one persistent resident wave with a static deal (profiles/r01_kernel_variants.md)
        k_msm_accumulate<<<gridfor(L.max_items, MSM_ACC_THREADS, (size_t)sm_count * 16), MSM_ACC_THREADS, 0, s>>>(
            bases, (unsigned)buckets_per_job, sorted, hist, offsets, item_off, item_bucket, n_items, L.chunk, order, partial);
    }
    if (ev) cudaEventRecord(ev[2], s);
    {
        static int coop_blocks_per_sm = 0;
        if (!coop_blocks_per_sm) {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&coop_blocks_per_sm, k_msm_merge_chunks, 128, 0) != cudaSuccess || coop_blocks_per_sm < 1) coop_blocks_per_sm = 1;
            if (coop_blocks_per_sm > 4) coop_blocks_per_sm = 4;
        }
        k_msm_merge_serial<<<gridfor(L.nbuckets, 128, (size_t)sm_count * 16), 128, 0, s>>>(partial, items, item_off, L.nbuckets, n_items);
        k_msm_merge_heavy<<<sm_count * 4, MSM_HEAVY_THREADS, 0, s>>>(partial, items, item_off, (const unsigned*)(scratch + L.heavy), n_items);
        void* args[] = {(void*)&partial, (void*)&items, (void*)&item_off, (void*)&item_bucket, (void*)&n_items};
        if ((e = cudaLaunchCooperativeKernel((const void*)k_msm_merge_chunks, dim3(sm_count * coop_blocks_per_sm), dim3(128), args, 0, s))) return e;
    }
    if (ev) cudaEventRecord(ev[3], s);
    {
        const size_t spans = (size_t)L.sets * L.blocks_per_window;
        g1_xyzz* spanR = blockres; g1_xyzz* spanL = blockres + spans; g1_xyzz* bits = (g1_xyzz*)(scratch + L.bits);
        k_msm_span_sums<<<gridfor(spans, 128, (size_t)sm_count * 16), 128, 0, s>>>(partial, items, item_off, nb, L.span, spans, spanR, spanL);
        // two blocks of the bit-sum kernel fit an SM (128 registers x 256 threads): slice every (set, slot) sum so that one wave covers the machine
        const unsigned slots = L.span_bits + 2, k0 = 0;
        unsigned P = (unsigned)(2 * (size_t)sm_count / (slots * L.sets));
        const unsigned maxP = L.blocks_per_window / (2 * MSM_SUM_THREADS);
        if (P > maxP) P = maxP;
        if (P > 32) P = 32;
        if (P < 1) P = 1;
        g1_xyzz* parts = bits;
        const size_t sum_smem = MSM_SUM_THREADS * sizeof(g1_xyzz) + (MSM_SUM_THREADS / 4) * QUAD_SLOTS * sizeof(fq_t);
        static bool smem_set = false;
        if (!smem_set) { if ((e = cudaFuncSetAttribute(k_msm_bit_sums, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sum_smem))) return e; smem_set = true; }
        k_msm_bit_sums<<<dim3(slots, L.sets, P), MSM_SUM_THREADS, sum_smem, s>>>(spanR, spanL, L.blocks_per_window, L.span_bits, k0, parts, g_msm_quad_tree);
        k_msm_bit_finish<<<dim3(slots, L.sets), MSM_FIN_THREADS, 0, s>>>(parts, P, L.span_bits, k0, wout, g_msm_quad_tree);
    }
    if (ev) cudaEventRecord(ev[4], s);
    if (launches) *launches += 18 + 2 * (unsigned)K;   // kernels only (the clear is a memset)
    return cudaGetLastError();
}

// ---------------------------------------------------------------- precomputed window tables: table[w * n + i] = 2^(c w) * bases[i]
// cur (XYZZ) <- 2^c * cur, one thread per point
__global__ void __launch_bounds__(128) k_msm_table_step(g1_xyzz* cur, size_t n, unsigned c) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        g1_xyzz p = ld_xyzz(cur + i);
        for (unsigned k = 0; k < c; ++k) p = p.dbl();
        st_xyzz(cur + i, p);
    }
}
__global__ void k_affine_to_xyzz(const g1_affine* in, size_t n, g1_xyzz* out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) st_xyzz(out + i, g1_xyzz::from_affine(ld_affine(in + i)));
}
// table: W * n affine points (window 0 = a copy of bases); cur: n XYZZ scratch points
cudaError_t msm_build_table(const g1_affine* bases, size_t n, unsigned c, g1_affine* table, g1_xyzz* cur, int sm_count, cudaStream_t s, unsigned* launches,
                            unsigned windows) {
    const unsigned W = windows ? windows : (255 + c - 1) / c;
    cudaError_t e;
    if ((e = cudaMemcpyAsync(table, bases, n * sizeof(g1_affine), cudaMemcpyDeviceToDevice, s))) return e;
    k_affine_to_xyzz<<<gridfor(n, 256, (size_t)sm_count * 8), 256, 0, s>>>(bases, n, cur);
    for (unsigned w = 1; w < W; ++w) {
        k_msm_table_step<<<gridfor(n, 128, (size_t)sm_count * 16), 128, 0, s>>>(cur, n, c);
        if ((e = launch_batch_to_affine(cur, n, table + (size_t)w * n, sm_count, s))) return e;
    }
    if (launches) *launches += 2 * W - 1;
    return cudaGetLastError();
}

cudaError_t launch_tau_powers(const fr_t& tau, size_t first, size_t n, fr_t* out, int sm_count, cudaStream_t s) {
    size_t chunks = (n + MSM_POW_SPAN - 1) / MSM_POW_SPAN;
    k_tau_powers<<<gridfor(chunks, 128, (size_t)sm_count * 8), 128, 0, s>>>(tau, first, n, out);
    return cudaGetLastError();
}
cudaError_t launch_fixed_base_mul(const fr_t* scalars, size_t n, const g1_affine* table, g1_xyzz* out, int sm_count, cudaStream_t s) {
    k_fixed_base_mul<<<gridfor(n, 128, (size_t)sm_count * 16), 128, 0, s>>>(scalars, n, table, out);
    return cudaGetLastError();
}
cudaError_t launch_batch_to_affine(const g1_xyzz* in, size_t n, g1_affine* out, int sm_count, cudaStream_t s) {
    size_t chunks = (n + MSM_INV_SPAN - 1) / MSM_INV_SPAN;
    k_batch_to_affine<<<gridfor(chunks, 128, (size_t)sm_count * 16), 128, 0, s>>>(in, n, out);
    return cudaGetLastError();
}
cudaError_t launch_jac_to_xyzz(const g1_jac* in, size_t n, g1_xyzz* out, int sm_count, cudaStream_t s) {
    k_jac_to_xyzz<<<gridfor(n, 128, (size_t)sm_count * 16), 128, 0, s>>>(in, n, out);
    return cudaGetLastError();
}
cudaError_t launch_affine_to_jac(const g1_affine* in, size_t n, g1_jac* out, int sm_count, cudaStream_t s) {
    k_affine_to_jac<<<gridfor(n, 256, (size_t)sm_count * 8), 256, 0, s>>>(in, n, out);
    return cudaGetLastError();
}

}  // namespace tsg
