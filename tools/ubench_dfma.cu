// tools/ubench_dfma.cu - would a double-precision (52-bit limb) multiplier beat the integer one on B200?
// VERDICT r1 item 4 asks for the measurement before any adoption.  The FP64 scheme (Emmart / Zheng / Weems, ARITH 2018): a 254-bit operand is
// 5 limbs of 52 bits held as doubles; one limb product needs its exact high and low halves,
//     hi = fma_rz(a, b, 2^104);   lo = fma_rz(a, b, (2^104 + 2^52) - hi);          // 2 DFMA + 1 DADD
// and the two halves are accumulated as 64-bit integers (bit patterns; the constants are subtracted in bulk): 2 x (IADD3 + IADD3.X).
// A Montgomery product = 25 limb products for a b + 25 for m p = 50 x (2 DFMA + 1 DADD + 4 integer adds), plus m_i = t_i n' mod 2^52 per limb.
// This program measures (1) the DFMA issue rate, (2) the rate of that instruction mix as an UPPER bound on "products per second" of such a
// multiplier (no carries resolved, no conversions - the real thing is slower), against 67 G products/s of the shipped IMAD multiplier
// (profiles/r01_kernel_variants.md).  One JSON line.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

// (1) independent DFMA chains: 8 accumulators per thread
__global__ void k_dfma(double* out, const double* in, int iters) {
    double a[8], acc[8];
    for (int i = 0; i < 8; ++i) { a[i] = in[i] + threadIdx.x; acc[i] = in[8 + i]; }
    const double b = in[16];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = __fma_rz(a[c], b, acc[c]);
    }
    double r = 0;
    for (int i = 0; i < 8; ++i) r += acc[i];
    if (r == 0.123) out[0] = r;
}

// (2) the instruction mix of one limb product, 4 independent limb products in flight per thread: 2 DFMA + 1 DADD + two 64-bit integer accumulations
__global__ void k_mix(unsigned long long* out, const double* in, int iters) {
    const double C1 = 20282409603651670423947251286016.0;                 // 2^104
    const double C2 = 20282409603651670423947251286016.0 + 4503599627370496.0;   // 2^104 + 2^52
    double a[4], b[4];
    unsigned long long sh[4], sl[4];
    for (int i = 0; i < 4; ++i) { a[i] = in[i] + threadIdx.x; b[i] = in[4 + i] + blockIdx.x; sh[i] = i; sl[i] = 3 * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double hi = __fma_rz(a[c], b[(c + u) & 3], C1);
                const double sub = C2 - hi;
                const double lo = __fma_rz(a[c], b[(c + u) & 3], sub);
                sh[c] += (unsigned long long)__double_as_longlong(hi);
                sl[c] += (unsigned long long)__double_as_longlong(lo);
            }
        }
        a[0] += 1.0; a[1] += 1.0; a[2] += 1.0; a[3] += 1.0;              // new operands every iteration: nothing can be hoisted
    }
    unsigned long long r = 0;
    for (int i = 0; i < 4; ++i) r ^= sh[i] ^ sl[i];
    if (r == 0x12345678ull) out[0] = r;
}

template <class F> float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double *d, h[32];
    CK(cudaMalloc(&d, 512));
    for (int i = 0; i < 32; ++i) h[i] = 1125899906842624.0 + 12345.0 * (i + 1);   // ~2^50: limb-sized integers
    CK(cudaMemcpy(d + 32, h, 256, cudaMemcpyHostToDevice));
    const int iters = 2000, blocks = sms * 8, threads = 256;
    float ms = time_ms([&] { k_dfma<<<blocks, threads>>>(d, d + 32, iters); }, 5);
    const double dfma = (double)blocks * threads * iters * 32.0 / (ms * 1e-3) / 1e12;
    ms = time_ms([&] { k_mix<<<blocks, threads>>>((unsigned long long*)d, d + 32, iters); }, 5);
    const double limb_products = (double)blocks * threads * iters * 16.0 / (ms * 1e-3);
    printf("{\"dfma_tops\": %.3f, \"limb_products_52x52_per_s_T\": %.3f, \"montgomery_products_upper_bound_G\": %.2f, \"shipped_imad_multiplier_G\": 67.0, "
           "\"note\": \"upper bound = limb-product mix rate / 50 (25 for a b + 25 for m p); carries, the m_i steps and the 32-bit <-> 52-bit conversions are not included\"}\n",
           dfma, limb_products / 1e12, limb_products / 50.0 / 1e9);
    return 0;
}
