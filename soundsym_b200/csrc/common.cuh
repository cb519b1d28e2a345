// common.cuh — context, error plumbing and small device helpers shared by the kernels of libsoundsym_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <algorithm>
#include <cstdio>
#include <map>
#include <mutex>
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
    std::vector<void*> pinned_free;  // 32-byte pinned blocks handed to dictionaries (the match's counters) and taken back
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

// Device-memory cache. A dictionary or query batch owns dozens of buffers; cudaMalloc / cudaFree cost 0.1 - 1 ms each (and
// cudaFree synchronises the device), so a flow that builds a fresh dictionary per match (examples/reconstruction.rs) paid
// 10 - 50 ms of allocator time per match. Blocks released by a handle's destructor AFTER its stream was synchronised
// (ss_dict_destroy, ss_queries_destroy: CacheFrees scope) go to a per-device free list instead of back to the driver: no work
// can still touch them, so any stream may take them. Blocks released while work may be pending (a workspace growing in the
// middle of a match) still go through cudaFree, which waits.
struct DevCache {
    static constexpr int kMaxDev = 16;
    static constexpr size_t kMaxBytes = 16ull << 30;  // cached per device; beyond that blocks go back to the driver
    std::mutex m;
    std::multimap<size_t, void*> blocks[kMaxDev];
    size_t bytes[kMaxDev] = {};
};
inline DevCache& dev_cache() {
    static DevCache* c = new DevCache();  // (never destroyed: handles may be released during process teardown)
    return *c;
}
inline thread_local bool tl_cache_frees = false;
struct CacheFrees {
    bool prev;
    CacheFrees() : prev(tl_cache_frees) { tl_cache_frees = true; }
    ~CacheFrees() { tl_cache_frees = prev; }
};
inline size_t dev_round(size_t bytes) {
    const size_t g = bytes < (1u << 20) ? 4096 : (1u << 20);
    return (bytes + g - 1) / g * g;
}
inline cudaError_t dev_alloc(void** p, size_t bytes, size_t* cap, int* dev) {
    *p = nullptr;
    cudaError_t e = cudaGetDevice(dev);
    if (e != cudaSuccess) return e;
    const size_t want = dev_round(bytes);
    DevCache& c = dev_cache();
    if (*dev >= 0 && *dev < DevCache::kMaxDev) {
        std::lock_guard<std::mutex> lock(c.m);
        auto it = c.blocks[*dev].lower_bound(want);
        if (it != c.blocks[*dev].end() && it->first <= want + std::max<size_t>(want / 4, 1u << 20)) {
            *p = it->second;
            *cap = it->first;
            c.bytes[*dev] -= it->first;
            c.blocks[*dev].erase(it);
            return cudaSuccess;
        }
    }
    e = cudaMalloc(p, want);
    if (e == cudaErrorMemoryAllocation && *dev >= 0 && *dev < DevCache::kMaxDev) {  // give the cached blocks back and try again
        cudaGetLastError();
        std::multimap<size_t, void*> drop;
        {
            std::lock_guard<std::mutex> lock(c.m);
            drop.swap(c.blocks[*dev]);
            c.bytes[*dev] = 0;
        }
        for (auto& b : drop) cudaFree(b.second);
        e = cudaMalloc(p, want);
    }
    *cap = want;
    return e;
}
inline void dev_free(void* p, size_t cap, int dev) {
    if (!p) return;
    if (tl_cache_frees && dev >= 0 && dev < DevCache::kMaxDev) {
        DevCache& c = dev_cache();
        std::lock_guard<std::mutex> lock(c.m);
        if (c.bytes[dev] + cap <= DevCache::kMaxBytes) {
            c.blocks[dev].emplace(cap, p);
            c.bytes[dev] += cap;
            return;
        }
    }
    cudaFree(p);
}

// RAII device buffer (allocation is synchronous; used at create time and for grow-only workspaces)
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    size_t cap_bytes = 0;
    int dev = -1;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) dev_free(p, cap_bytes, dev);
        p = nullptr;
        n = 0;
    }
    // grow-only
    cudaError_t reserve(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        if (count == 0) count = 1;
        void* q = nullptr;
        cudaError_t e = dev_alloc(&q, count * sizeof(T), &cap_bytes, &dev);
        if (e == cudaSuccess) p = static_cast<T*>(q), n = count;
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
