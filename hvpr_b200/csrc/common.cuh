// Shared helpers for the hvpr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/hvpr_b200.h"

namespace hvpr {

constexpr int kMaxSMs = 148;  // B200; sizes host-side workspace queries that must work without a device
// SM count of the CURRENT device, queried once per device and cached (immutable, so not "state"); kMaxSMs when no device answers
int num_sms();

void set_cuda_error(cudaError_t e);

#define HVPR_CHECK_LAUNCH()                                         \
    do {                                                            \
        cudaError_t e__ = cudaGetLastError();                       \
        if (e__ != cudaSuccess) { hvpr::set_cuda_error(e__); return HVPR_ERR_CUDA; } \
    } while (0)

#define HVPR_CHECK_CUDA(call)                                       \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) { hvpr::set_cuda_error(e__); return HVPR_ERR_CUDA; } \
    } while (0)

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// streaming (evict-first) 128-bit store: canvases are written once and not re-read by this path
#ifndef HVPR_BEV_STORE
#define HVPR_BEV_STORE 0
#endif
__device__ __forceinline__ void st_stream_f4(float4 *p, float4 v) {
#if HVPR_BEV_STORE == 0
    __stcs(p, v);          // evict-first
#elif HVPR_BEV_STORE == 1
    *p = v;                // default write-back
#else
    __stwt(p, v);          // write-through
#endif
}

}  // namespace hvpr
