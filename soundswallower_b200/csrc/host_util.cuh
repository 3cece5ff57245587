// host_util.cuh -- host-side helpers shared by the C-ABI translation units:
// CUDA error -> ssb_last_error plumbing and a grow-only device buffer.
#pragma once
#include <cuda_runtime.h>

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

// grow-only device buffer
struct DBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap)
            return 0;
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&p, bytes);
            want = bytes;
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
        if (p)
            cudaFree(p);
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
