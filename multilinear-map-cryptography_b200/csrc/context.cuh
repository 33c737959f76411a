// context.cuh - internal definitions behind the opaque handles of include/tsgpu.h
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include "../../include/tsgpu.h"
#include "fp.cuh"
#include "sumcheck.cuh"

struct tsgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    std::string err;
    uint64_t launches = 0;
    // small persistent scratch: block partial sums, ticket counter, 8 result slots, pinned mirror
    tsg::fr_t* partials = nullptr;
    unsigned int* ticket = nullptr;
    tsg::fr_t* dev_out = nullptr;      // 8 elements
    tsg::fr_t* host_out = nullptr;     // pinned, 8 elements
    void* comm = nullptr;              // multi-GPU communicator (comm.cu), optional
    void* interp = nullptr;            // cached interpolation plan (interp.cu)
};

struct tsgpu_table {
    tsg::fr_t* d = nullptr;   // 2^num_vars entries, bit-reversed index order
    unsigned num_vars = 0;
    size_t capacity = 0;      // allocated entries
};

struct tsgpu_sc {
    tsgpu_ctx* ctx = nullptr;
    int d = 0;
    unsigned vars_left = 0;
    tsgpu_table* tables[tsg::SC_MAX_TABLES] = {nullptr, nullptr, nullptr};
};

namespace tsg {

int fail(tsgpu_ctx* ctx, int code, const std::string& msg);
int cuda_fail(tsgpu_ctx* ctx, cudaError_t e, const char* what);

#define TSG_CUDA(ctx, call)                                                   \
    do {                                                                      \
        cudaError_t _e = (call);                                              \
        if (_e != cudaSuccess) return ::tsg::cuda_fail((ctx), _e, #call);     \
    } while (0)

// stream-ordered temporary device buffer
struct TempBuf {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    cudaError_t alloc(size_t bytes, cudaStream_t stream) {
        s = stream;
        return cudaMallocAsync(&p, bytes ? bytes : 32, stream);
    }
    ~TempBuf() { if (p) cudaFreeAsync(p, s); }
    template <class T> T* as() { return (T*)p; }
};

int table_alloc(tsgpu_ctx* ctx, unsigned num_vars, tsgpu_table** out);

}  // namespace tsg
