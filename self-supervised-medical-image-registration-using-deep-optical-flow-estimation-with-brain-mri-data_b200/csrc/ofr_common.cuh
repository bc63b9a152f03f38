// Shared host/device helpers for the ofr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>

#include "../../include/ofr_b200.h"

namespace ofr {

extern std::atomic<uint64_t> g_launches;  // instrumentation only (ofr_launch_count)

inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? OFR_OK : OFR_ECUDA + (int)e; }

// Call after every launch: counts it and turns a launch error into a return code.
inline int after_launch() {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cuda_rc(cudaPeekAtLastError());
}

#define OFR_LAUNCH_CHECK()                 \
    do {                                   \
        int _rc = ::ofr::after_launch();   \
        if (_rc != OFR_OK) return _rc;     \
    } while (0)

#define OFR_CUDA(expr)                                  \
    do {                                                \
        cudaError_t _e = (expr);                        \
        if (_e != cudaSuccess) return OFR_ECUDA + (int)_e; \
    } while (0)

inline int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    return n;
}

// Grid for a grid-stride kernel: enough CTAs for `work` items, capped at a whole
// number of waves (ctas_per_sm resident CTAs on each of the B200's 148 SMs).
inline unsigned grid_for(long long work, int threads, int ctas_per_sm) {
    long long need = (work + threads - 1) / threads;
    long long cap = (long long)sm_count() * ctas_per_sm;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming (evict-first) 128-bit store for write-once outputs
__device__ __forceinline__ void stcs4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

}  // namespace ofr
