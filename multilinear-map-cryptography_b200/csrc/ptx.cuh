// ptx.cuh - carry-chain integer primitives for 256-bit modular arithmetic on sm_100a.
//
// Each primitive is one PTX instruction on the device (mad.lo.cc / madc.hi.cc pairs are fused by
// ptxas into IMAD.WIDE.U32 with a predicate carry, which is what keeps a Montgomery product at
// ~2n^2+n integer-pipe instructions).  The same functions have a host emulation (explicit carry
// flag) so the limb algorithms in fp.cuh can be unit-tested bit-for-bit on a CPU-only box
// (tests/test_limb_arith_host.py).  The host side is also what the C++ host mirror (host/*.cpp) uses
// for the handful of field operations that stay on the CPU: transcript serialisation, 4-point
// interpolation and Horner evaluation of round polynomials, final MSM window combination.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define TSG_HD __host__ __device__ __forceinline__
#define TSG_D __device__ __forceinline__
#else
#define TSG_HD inline
#define TSG_D inline
#endif

namespace tsg {
namespace ptx {

#if !defined(__CUDA_ARCH__)
// host emulation of the PTX condition-code carry flag
inline uint32_t& cf() { static thread_local uint32_t f = 0; return f; }
inline uint32_t emu_add(uint32_t a, uint32_t b, uint32_t cin, bool set) {
    uint64_t s = (uint64_t)a + b + cin;
    if (set) cf() = (uint32_t)(s >> 32);
    return (uint32_t)s;
}
inline uint32_t emu_sub(uint32_t a, uint32_t b, uint32_t bin, bool set) {
    uint64_t s = (uint64_t)a - b - bin;
    if (set) cf() = (uint32_t)((s >> 32) & 1);   // borrow
    return (uint32_t)s;
}
inline uint32_t lo32(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t hi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
#endif

TSG_HD uint32_t add_cc(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return emu_add(a, b, 0, true);
#endif
}
TSG_HD uint32_t addc_cc(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return emu_add(a, b, cf(), true);
#endif
}
TSG_HD uint32_t addc(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return emu_add(a, b, cf(), false);
#endif
}
TSG_HD uint32_t sub_cc(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return emu_sub(a, b, 0, true);
#endif
}
TSG_HD uint32_t subc_cc(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return emu_sub(a, b, cf(), true);
#endif
}
TSG_HD uint32_t subc(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return emu_sub(a, b, cf(), false);
#endif
}
TSG_HD uint32_t mul_lo(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return lo32(a, b);
#endif
}
TSG_HD uint32_t mul_hi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return hi32(a, b);
#endif
}
// full 32x32 -> 64 product as one IMAD.WIDE (no carry predicate: full issue rate)
TSG_HD void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    unsigned long long r;
    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    lo = (uint32_t)r; hi = (uint32_t)(r >> 32);
#else
    unsigned long long r = (unsigned long long)a * b;
    lo = (uint32_t)r; hi = (uint32_t)(r >> 32);
#endif
}
TSG_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    return emu_add(lo32(a, b), c, 0, true);
#endif
}
TSG_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    return emu_add(lo32(a, b), c, cf(), true);
#endif
}
TSG_HD uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    return emu_add(hi32(a, b), c, 0, true);
#endif
}
TSG_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    return emu_add(hi32(a, b), c, cf(), true);
#endif
}
TSG_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    return emu_add(hi32(a, b), c, cf(), false);
#endif
}

}  // namespace ptx
}  // namespace tsg
