// irb_common.hpp -- host-side helpers shared by the translation units of libirb_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstddef>

#include "../../include/irb_b200.h"

namespace irbh {

int fail(int code, const char* fmt, ...);                      // records the thread's last error, returns code
int current_device();                                          // device chosen by irb_set_device on this thread
int twiddles(int dev, int M, const float2** out);              // table of the 2M roots exp(-2 pi i k / 2M), cached per (device, M)
extern std::atomic<long long> g_launches;

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return irbh::fail(IRB_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes, bool zero) {
        CK(cudaMalloc(&p, bytes ? bytes : 16));
        if (zero) CK(cudaMemset(p, 0, bytes ? bytes : 16));
        return 0;
    }
    template <typename T> T* as() const { return (T*) p; }
};

struct StreamGuard {
    cudaStream_t s = nullptr;
    ~StreamGuard() { if (s) cudaStreamDestroy(s); }
    int create() { CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); return 0; }
};

inline int next_pow2(int x) {                                  // tools::nextPowerOfTwo, fp/tools.cpp:189-196
    if (x > 0 && (x & (x - 1)) == 0) return x;
    int r = 1;
    while (r <= x) r *= 2;
    return r;
}

}  // namespace irbh
