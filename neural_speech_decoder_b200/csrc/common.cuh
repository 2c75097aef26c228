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

// Philox4x32-10 counter-based generator: 4 uniform 32-bit words per (key, counter).
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}

// Inter-layer dropout (reference model.py:55): the keep/drop decision of element e of a contiguous tensor depends only on
// (seed, e): one Philox call yields the bits for elements 4q .. 4q+3; element i is KEPT iff word i >= dropout_threshold(p).
__device__ __forceinline__ uint4 dropout_bits(size_t q, uint64_t seed) {
    return philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) { return (uint32_t)fminf(4294967295.0f, p * 4294967296.0f); }

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

}  // namespace nsd
