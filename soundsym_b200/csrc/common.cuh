// common.cuh — context, error plumbing and small device helpers shared by the kernels of libsoundsym_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <new>
#include <string>
#include <vector>

#include "../../include/soundsym_b200.h"

struct ss_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    std::string err;
    uint64_t launches = 0;
    // per-subsystem state (tables, grow-only workspaces) owned by the ctx; created lazily by sound.cu / segment.cu
    void* sound_state = nullptr;
    void (*sound_state_free)(void*) = nullptr;
    void* seg_state = nullptr;
    void (*seg_state_free)(void*) = nullptr;
};

namespace ss {

int set_error(ss_ctx* ctx, int code, const char* fmt, ...);

#define SS_CUDA(ctx, expr)                                                                                   \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess)                                                                               \
            return ss::set_error((ctx), _e == cudaErrorMemoryAllocation ? SS_ERR_NOMEM : SS_ERR_CUDA,       \
                                 "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define SS_TRY(expr)             \
    do {                         \
        int _rc = (expr);        \
        if (_rc != SS_OK) return _rc; \
    } while (0)

// counts one launch of one of this library's kernels and checks the launch
#define SS_LAUNCHED(ctx)                    \
    do {                                    \
        (ctx)->launches++;                  \
        SS_CUDA((ctx), cudaGetLastError()); \
    } while (0)

// RAII device buffer bound to a ctx's device (allocation is synchronous; used at create time and for workspaces)
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    // grow-only
    cudaError_t reserve(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        else p = nullptr;
        return e;
    }
};

template <typename T>
inline int upload(ss_ctx* ctx, DevBuf<T>& dst, const T* src, size_t count) {
    SS_CUDA(ctx, dst.reserve(count));
    if (count) SS_CUDA(ctx, cudaMemcpyAsync(dst.p, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return SS_OK;
}

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace ss
