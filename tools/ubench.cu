// tools/ubench.cu - integer-pipe and HBM micro-benchmarks for B200 (sm_100a).
// Measures the denominators DESIGN.md quotes for integer-bound kernels: IMAD / IMAD.WIDE / IADD3 issue
// rates, Montgomery products per second (Fr), and a 256-bit-load streaming copy.  Prints one JSON object.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../multilinear-map-cryptography_b200/csrc/fr_device.cuh"
using namespace tsg;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

template <int MODE>
__global__ void k_int(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t x0 = threadIdx.x, x1 = a, x2 = b, x3 = a ^ b, x4 = 5, x5 = 7, x6 = 11, x7 = 13;
    uint64_t w0 = x0, w1 = x1, w2 = x2, w3 = x3, w4 = 1, w5 = 2, w6 = 3, w7 = 4;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (MODE == 0) {   // IMAD (32-bit)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x0) : "r"(a), "r"(b));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x1) : "r"(a), "r"(b));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x2) : "r"(a), "r"(b));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x3) : "r"(a), "r"(b));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x4) : "r"(a), "r"(b));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x5) : "r"(a), "r"(b));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x6) : "r"(a), "r"(b));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x7) : "r"(a), "r"(b));
            } else if (MODE == 1) {   // IMAD.WIDE.U32
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w0) : "r"(a), "r"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w1) : "r"(a), "r"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w2) : "r"(a), "r"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w3) : "r"(a), "r"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w4) : "r"(a), "r"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w5) : "r"(a), "r"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w6) : "r"(a), "r"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w7) : "r"(a), "r"(b));
            } else if (MODE == 2) {   // IMAD.HI.U32
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x0) : "r"(a), "r"(b));
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x1) : "r"(a), "r"(b));
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x2) : "r"(a), "r"(b));
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x3) : "r"(a), "r"(b));
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x4) : "r"(a), "r"(b));
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x5) : "r"(a), "r"(b));
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x6) : "r"(a), "r"(b));
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x7) : "r"(a), "r"(b));
            } else {   // IADD3-class
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x0) : "r"(a));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x1) : "r"(b));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x2) : "r"(a));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x3) : "r"(b));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x4) : "r"(a));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x5) : "r"(b));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x6) : "r"(a));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x7) : "r"(b));
            }
        }
    }
    uint32_t r = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7 ^ (uint32_t)(w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7) ^ (uint32_t)((w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7) >> 32);
    if (r == 0x12345678u) out[0] = r;
}

__global__ void k_montmul(fr_t* out, const fr_t* in, int iters) {
    fr_t a = in[threadIdx.x & 7], b = in[8 + (threadIdx.x & 7)];
    fr_t c = a, d = b;
    for (int i = 0; i < iters; ++i) { c = c * a; d = d * b; }
    fr_t r = c + d;
    if (r.l[0] == 0x12345678u && r.l[7] == 1) out[0] = r;
}
__global__ void k_wide(fr_t* out, const fr_t* in, int iters) {
    fr_t a = in[threadIdx.x & 7], b = in[8 + (threadIdx.x & 7)];
    wide_acc<FrP> acc; acc.clear();
    for (int i = 0; i < iters; ++i) { acc.add_product(a, b); a.l[0] ^= acc.t[3]; }
    fr_t r = acc.reduce();
    if (r.l[0] == 0x12345678u && r.l[7] == 1) out[0] = r;
}
__global__ void k_copy256(const fr_t* in, fr_t* out, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) st256(out + i, ld256_stream(in + i));
}
__global__ void k_read256(const fr_t* in, fr_t* out, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) { fr_t v = ld256_stream(in + i); acc ^= v.l[0] ^ v.l[7]; }
    if (acc == 0x12345678u) out[0].l[0] = acc;
}

template <class F> float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f(); f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    uint32_t* d32; CK(cudaMalloc(&d32, 64));
    fr_t* dfr; CK(cudaMalloc(&dfr, 64 * sizeof(fr_t)));
    fr_t h[16]; for (int i = 0; i < 16; ++i) for (int k = 0; k < 8; ++k) h[i].l[k] = 0x01234567u * (i + 1) + k * 0x9e3779b9u; for (int i = 0; i < 16; ++i) h[i].l[7] &= 0x0fffffffu;
    CK(cudaMemcpy(dfr + 16, h, sizeof(h), cudaMemcpyHostToDevice));
    const int iters = 2000, blocks = sms * 8, threads = 256;
    double nthreads = (double)blocks * threads;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    const char* names[4] = {"imad32", "imad_wide", "imad_hi", "iadd"};
    for (int m = 0; m < 4; ++m) {
        float ms = 0;
        if (m == 0) ms = time_ms([&] { k_int<0><<<blocks, threads>>>(d32, 3, 5, iters); }, 5);
        if (m == 1) ms = time_ms([&] { k_int<1><<<blocks, threads>>>(d32, 3, 5, iters); }, 5);
        if (m == 2) ms = time_ms([&] { k_int<2><<<blocks, threads>>>(d32, 3, 5, iters); }, 5);
        if (m == 3) ms = time_ms([&] { k_int<3><<<blocks, threads>>>(d32, 3, 5, iters); }, 5);
        double ops = nthreads * iters * 16.0 * 8.0;
        printf(", \"%s_tops\": %.3f", names[m], ops / (ms * 1e-3) / 1e12);
    }
    {
        const int it = 500;
        float ms = time_ms([&] { k_montmul<<<blocks, threads>>>(dfr, dfr + 16, it); }, 5);
        printf(", \"fr_montmul_gops\": %.2f", nthreads * it * 2.0 / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { k_wide<<<blocks, threads>>>(dfr, dfr + 16, it); }, 5);
        printf(", \"fr_wide_mac_gops\": %.2f", nthreads * it / (ms * 1e-3) / 1e9);
    }
    {
        size_t n = (size_t)1 << 26;   // 2 GiB per buffer
        fr_t *a, *b; CK(cudaMalloc(&a, n * 32)); CK(cudaMalloc(&b, n * 32));
        CK(cudaMemset(a, 1, n * 32));
        float ms = time_ms([&] { k_copy256<<<sms * 16, 256>>>(a, b, n); }, 10);
        printf(", \"copy256_gbs\": %.1f", 2.0 * n * 32 / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { k_read256<<<sms * 16, 256>>>(a, b, n); }, 10);
        printf(", \"read256_gbs\": %.1f", 1.0 * n * 32 / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { cudaMemcpyAsync(b, a, n * 32, cudaMemcpyDeviceToDevice); }, 10);
        printf(", \"memcpy_d2d_gbs\": %.1f", 2.0 * n * 32 / (ms * 1e-3) / 1e9);
        cudaFree(a); cudaFree(b);
    }
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf(", \"sm_clock_khz_attr\": %d}\n", clk);
    return 0;
}
