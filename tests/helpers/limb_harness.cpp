// Host-emulation harness for the limb algorithms in csrc/fp.cuh (carry flag emulated in ptx.cuh).
// Built by tests/test_limb_arith_host.py with g++; lets the exact device algorithms be checked
// against the oracle without a GPU.  Not part of the product.
#include <cstddef>
#include <cstring>
#include "fp.cuh"
using namespace tsg;

template <class P>
static void binop(int op, const uint32_t* a, const uint32_t* b, size_t n, uint32_t* out) {
    for (size_t i = 0; i < n; ++i) {
        fp<P> x, y, r;
        memcpy(x.l, a + 8 * i, 32); memcpy(y.l, b + 8 * i, 32);
        switch (op) {
            case 0: r = x + y; break;
            case 1: r = x - y; break;
            case 2: r = x * y; break;
            case 3: { uint32_t t[16]; limb::mul_wide(t, x.l, y.l); limb::wide_normalize<P>(t); limb::mont_reduce<P>(r.l, t); break; }
            case 4: r = x.from_mont(); break;
            case 5: r = x.to_mont(); break;
            case 6: r = x.inverse(); break;
            case 7: r = x.neg(); break;
            case 8: limb::mont_mul<P>(r.l, x.l, y.l); break;
            case 10: r = fp_ctab<P>::make(y).mul(x); break;       // constant-multiplier table of y applied to x
            case 11: r = x.sqr(); break;
            default: r = fp<P>::zero();
        }
        memcpy(out + 8 * i, r.l, 32);
    }
}
extern "C" void limb_binop(int field, int op, const uint32_t* a, const uint32_t* b, size_t n, uint32_t* out) {
    if (field == 0) binop<FrP>(op, a, b, n, out); else binop<FqP>(op, a, b, n, out);
}
// sum_i a_i * b_i through the lazy 512-bit accumulator
extern "C" void limb_dot(int field, const uint32_t* a, const uint32_t* b, size_t n, uint32_t* out) {
    if (field == 0) {
        wide_acc<FrP> acc; acc.clear();
        for (size_t i = 0; i < n; ++i) { fr_t x, y; memcpy(x.l, a + 8 * i, 32); memcpy(y.l, b + 8 * i, 32); acc.add_product(x, y); }
        fr_t r = acc.reduce(); memcpy(out, r.l, 32);
    } else {
        wide_acc<FqP> acc; acc.clear();
        for (size_t i = 0; i < n; ++i) { fq_t x, y; memcpy(x.l, a + 8 * i, 32); memcpy(y.l, b + 8 * i, 32); acc.add_product(x, y); }
        fq_t r = acc.reduce(); memcpy(out, r.l, 32);
    }
}
// r_i = a_i b_i - c_i d_i with one reduction (limb::mont_mul_sub)
extern "C" void limb_mul_sub(int field, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, size_t n, uint32_t* out) {
    for (size_t i = 0; i < n; ++i) {
        if (field == 0) limb::mont_mul_sub<FrP>(out + 8 * i, a + 8 * i, b + 8 * i, c + 8 * i, d + 8 * i);
        else limb::mont_mul_sub<FqP>(out + 8 * i, a + 8 * i, b + 8 * i, c + 8 * i, d + 8 * i);
    }
}
extern "C" void limb_mul_wide(const uint32_t* a, const uint32_t* b, uint32_t* out16) { limb::mul_wide(out16, a, b); }
