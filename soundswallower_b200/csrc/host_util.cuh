// host_util.cuh -- host-side helpers shared by the C-ABI translation units:
// CUDA error -> ssb_last_error plumbing and a grow-only device buffer.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <map>
#include <mutex>

#include "model.h"

#define API_CUDA(call, rv)                                                                  \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            ssb::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,             \
                           cudaGetErrorString(e_));                                         \
            return rv;                                                                      \
        }                                                                                   \
    } while (0)

namespace ssb {

// Device blocks released by one call are kept for the next one: the one-shot entry points
// (ssb_align_batch, ssb_fsg_batch, ssb_align_texts) allocate 10+ GB of intermediates per call,
// and cudaMalloc / cudaFree of that costs tens of milliseconds.  Per device, capped at
// $SSB_DEV_CACHE_GB (default 64, 0 = off); ssb_device_cache_trim() gives everything back, and
// a failed cudaMalloc trims the cache before it gives up.
struct DevBlockCache {
    std::mutex mu;
    std::multimap<size_t, void *> free_[16];
    size_t held[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    size_t cap;
    DevBlockCache()
    {
        const char *e = getenv("SSB_DEV_CACHE_GB");
        cap = (size_t)(e ? atof(e) : 64.0) << 30;
    }
    void trim(int dev)
    {
        for (auto &kv : free_[dev])
            cudaFree(kv.second);
        free_[dev].clear();
        held[dev] = 0;
    }
};
inline DevBlockCache &dev_block_cache()
{
    static DevBlockCache c;
    return c;
}
inline cudaError_t cached_malloc(void **p, size_t bytes, size_t *got)
{
    DevBlockCache &c = dev_block_cache();
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 15;
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.free_[dev].lower_bound(bytes);
        if (it != c.free_[dev].end() && it->first <= bytes + bytes / 2 + (1u << 20)) {
            *p = it->second;
            *got = it->first;
            c.held[dev] -= it->first;
            c.free_[dev].erase(it);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        std::lock_guard<std::mutex> lk(c.mu);
        c.trim(dev);
        e = cudaMalloc(p, bytes);
    }
    *got = bytes;
    return e;
}
inline void cached_free(void *p, size_t bytes)
{
    if (!p)
        return;
    DevBlockCache &c = dev_block_cache();
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 15;
    std::lock_guard<std::mutex> lk(c.mu);
    if (bytes < (1u << 16) || c.held[dev] + bytes > c.cap) {
        cudaFree(p);
        return;
    }
    c.free_[dev].insert({bytes, p});
    c.held[dev] += bytes;
}

// grow-only device buffer
struct DBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap)
            return 0;
        cached_free(p, cap);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cached_malloc(&p, want, &want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cached_malloc(&p, bytes, &want);
        }
        if (e != cudaSuccess) {
            ssb::set_error("cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
            p = nullptr;
            return -1;
        }
        cap = want;
        return 0;
    }
    void release()
    {
        cached_free(p, cap);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const
    {
        return reinterpret_cast<T *>(p);
    }
};

}  // namespace ssb
