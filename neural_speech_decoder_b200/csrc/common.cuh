// Shared helpers for the nsd_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nsd_b200.h"

namespace nsd {

void set_error(const char* fmt, ...);

#define NSD_CHECK_ARG(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            nsd::set_error(__VA_ARGS__);    \
            return NSD_ERR_INVALID;         \
        }                                   \
    } while (0)

#define NSD_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            nsd::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return NSD_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

void count_launch(int n);   // kernels enqueued by this library (nsd_launch_count)

#define NSD_LAUNCH_CHECK()              \
    do {                                \
        nsd::count_launch(1);           \
        NSD_CUDA(cudaGetLastError());   \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t cdivz(size_t a, size_t b) { return (a + b - 1) / b; }

int sm_count();   // cached multiProcessorCount of the current device

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
// tanh via exp, accurate to ~1e-7 relative for the fp32 parity path
__device__ __forceinline__ float tanhf_(float x) { return tanhf(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

}  // namespace nsd
