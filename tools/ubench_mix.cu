// tools/ubench_mix.cu - Montgomery products per second for the multiplier variants selected at compile time
// (-DTSG_MIX_ROWS=k: k word steps use plain IMAD.WIDE + ALU additions for their reduction rows).  Also times a mixed-addition
// loop (the MSM inner operation).  Prints one JSON object.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../multilinear-map-cryptography_b200/csrc/fr_device.cuh"
#include "../multilinear-map-cryptography_b200/csrc/g1.cuh"
using namespace tsg;

__global__ void k_montmul(fr_t* out, const fr_t* in, int iters) {
    fr_t a = in[threadIdx.x & 7], b = in[8 + (threadIdx.x & 7)];
    fr_t c = a, d = b;
    for (int i = 0; i < iters; ++i) { c = c * a; d = d * b; }
    fr_t r = c + d;
    if (r.l[0] == 0x12345678u && r.l[7] == 1) out[0] = r;
}
__global__ void __launch_bounds__(128) k_madd(fq_t* out, const fq_t* in, int iters) {
    g1_affine p; p.x = in[threadIdx.x & 7]; p.y = in[8 + (threadIdx.x & 7)];
    g1_xyzz acc; acc.X = in[(threadIdx.x + 1) & 7]; acc.Y = in[(threadIdx.x + 3) & 7]; acc.ZZ = in[(threadIdx.x + 5) & 7]; acc.ZZZ = in[(threadIdx.x + 2) & 7];
    for (int i = 0; i < iters; ++i) { acc = acc.add_affine(p, i & 1); p.x.l[0] ^= acc.X.l[1]; }
    if (acc.X.l[0] == 0x12345678u && acc.Y.l[7] == 1) out[0] = acc.X;
}
template <class F> float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    fr_t* dfr; cudaMalloc(&dfr, 64 * sizeof(fr_t));
    fr_t h[16]; for (int i = 0; i < 16; ++i) for (int k = 0; k < 8; ++k) h[i].l[k] = 0x01234567u * (i + 1) + k * 0x9e3779b9u; for (int i = 0; i < 16; ++i) h[i].l[7] &= 0x0fffffffu;
    cudaMemcpy(dfr + 16, h, sizeof(h), cudaMemcpyHostToDevice);
    const int blocks = sms * 8, threads = 256, it = 500;
    double nthreads = (double)blocks * threads;
    float ms = time_ms([&] { k_montmul<<<blocks, threads>>>(dfr, dfr + 16, it); }, 5);
    printf("{\"mix_rows\": %d, \"fr_montmul_gops\": %.2f", TSG_MIX_ROWS, nthreads * it * 2.0 / (ms * 1e-3) / 1e9);
    const int b2 = sms * 4, t2 = 128, it2 = 300;
    ms = time_ms([&] { k_madd<<<b2, t2>>>((fq_t*)dfr, (fq_t*)(dfr + 16), it2); }, 5);
    printf(", \"mixed_add_gops\": %.3f}\n", (double)b2 * t2 * it2 / (ms * 1e-3) / 1e9);
    return 0;
}
