// Element-wise / normalisation / depthwise-convolution kernels of the Conformer path (BASELINE configs[2]; reference
// src/neural_decoder/transformer_ctc.py).  All HBM-bound: one pass over the activation per stage, 16-byte accesses, the
// stochastic regularisers (nn.Dropout, DropPath) fused into the producing kernel as counter-based masks (Philox keyed by
// (seed, element) / (seed, sample): the backward regenerates the mask instead of storing it).
//   LayerNorm (+SiLU | GELU) (+dropout)      transformer_ctc.py:97-98, 156, 167, 202, 212, 219, 231, 410-413
//   SiLU / ReLU / GELU (+dropout)            transformer_ctc.py:140, 204-205, 223-224
//   GLU                                      transformer_ctc.py:160, 179
//   depthwise conv k (pad k/2) over time     transformer_ctc.py:162-166, 181-184; Gaussian smoothing :104-109
//   strided depthwise conv (k32/s4, no pad)  transformer_ctc.py:81-91, 112-114
//   x + scale * DropPath(dropout(y))         transformer_ctc.py:14-23, 245, 251, 190, 257
//   SpecAugment bands + positional encoding  transformer_ctc.py:266-308, 311-330, 467-471
#include <algorithm>

#include "common.cuh"

namespace nsd {

enum { ACT_NONE = 0, ACT_SILU = 1, ACT_GELU = 2, ACT_RELU = 3 };

__device__ __forceinline__ float act_fwd(float v, int act) {
    if (act == ACT_SILU) return v / (1.0f + __expf(-v));
    if (act == ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));          // nn.GELU() default: exact erf form
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    return v;
}
__device__ __forceinline__ float act_grad(float v, int act) {
    if (act == ACT_SILU) { const float s = 1.0f / (1.0f + __expf(-v)); return s * (1.0f + v * (1.0f - s)); }
    if (act == ACT_GELU) return 0.5f * (1.0f + erff(v * 0.70710678118654752f)) + v * 0.39894228040143268f * __expf(-0.5f * v * v);
    if (act == ACT_RELU) return v > 0.f ? 1.f : 0.f;
    return 1.f;
}
// keep/scale factor of element e under dropout(p): the word (e & 3) of the Philox block of e >> 2 (same scheme as nsd_dropout)
__device__ __forceinline__ float4 drop4(size_t q, float p, float inv_keep, uint64_t seed) {
    if (p <= 0.f) return make_float4(1.f, 1.f, 1.f, 1.f);
    const uint4 r = dropout_bits(q, seed);
    const uint32_t th = dropout_threshold(p);
    return make_float4(r.x >= th ? inv_keep : 0.f, r.y >= th ? inv_keep : 0.f, r.z >= th ? inv_keep : 0.f, r.w >= th ? inv_keep : 0.f);
}
// DropPath (transformer_ctc.py:14-23): per sample, keep with probability 1-p and scale by 1/(1-p)
__device__ __forceinline__ float path_factor(int b, float p, uint64_t seed) {
    if (p <= 0.f) return 1.f;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)b, 0u, 0x70617468u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    return r.x >= dropout_threshold(p) ? 1.0f / (1.0f - p) : 0.f;
}
__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* p, float4 v) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}
__device__ __forceinline__ float4 load4(const void* p, int dtype, size_t i) {       // 4 consecutive elements starting at i (i % 4 == 0)
    if (dtype == NSD_F32) return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p) + i);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

// ------------------------------------------------------------------------------------------------ LayerNorm (+act) (+dropout)
// one warp per row; the row stays in registers between the statistics and the normalisation (D <= 32 * 4 * LN_MAXV)
constexpr int LN_MAXV = 16;     // float4s per lane: D <= 2048 (kernels are instantiated for 4, 8 and 16)

template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float eps, int act, float p, uint64_t seed, const unsigned long long* __restrict__ seed_off,
                                                            float* __restrict__ y32, __nv_bfloat16* __restrict__ y16, float* __restrict__ mean, float* __restrict__ rstd,
                                                            int M, int D) {
    pdl_enter();
    if (seed_off) seed += *seed_off;
    const int lane = threadIdx.x & 31, row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int nv = D >> 2;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
    float4 v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < nv) { v[i] = xr[c]; s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
    }
    const float mu = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < nv) { const float a = v[i].x - mu, b = v[i].y - mu, cc = v[i].z - mu, d = v[i].w - mu; q += (a * a + b * b) + (cc * cc + d * d); }
    }
    const float rs = rsqrtf(warp_sum(q) / (float)D + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
    const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < nv) {
            const float4 g = reinterpret_cast<const float4*>(gamma)[c], b = reinterpret_cast<const float4*>(beta)[c];
            const float4 k = drop4(((size_t)row * D >> 2) + c, p, inv_keep, seed);
            float4 o;
            o.x = act_fwd(fmaf((v[i].x - mu) * rs, g.x, b.x), act) * k.x; o.y = act_fwd(fmaf((v[i].y - mu) * rs, g.y, b.y), act) * k.y;
            o.z = act_fwd(fmaf((v[i].z - mu) * rs, g.z, b.z), act) * k.z; o.w = act_fwd(fmaf((v[i].w - mu) * rs, g.w, b.w), act) * k.w;
            if (y32) reinterpret_cast<float4*>(y32 + (size_t)row * D)[c] = o;
            if (y16) store_bf16x4(y16 + (size_t)row * D + 4 * c, o);
        }
    }
}

// backward in two light kernels instead of one heavy one (a fused form that also carried the dgamma / dbeta column sums needed 120 registers
// and 64 KB of shared memory per CTA: 16 warps per SM, long-scoreboard-bound at 20 % of DRAM peak under ncu):
//   layernorm_bwd_dx_kernel     warp per row, row in registers, no column state -> 4 CTAs per SM;
//   layernorm_bwd_param_kernel  thread = 4 columns x every 8th row of a 256-row chunk (re-reads dy and x: +1 pass, coalesced, deep MLP),
//                               per-chunk partial sums reduced in fixed order by layernorm_param_reduce_kernel: deterministic.
__device__ __forceinline__ float4 ln_bwd_dterm(float4 d, const float4& xh, const float4& g, const float* __restrict__ beta, int c, int act, const float4& k) {
    d.x *= k.x; d.y *= k.y; d.z *= k.z; d.w *= k.w;
    if (act != ACT_NONE) {
        const float4 b = reinterpret_cast<const float4*>(beta)[c];
        d.x *= act_grad(fmaf(xh.x, g.x, b.x), act); d.y *= act_grad(fmaf(xh.y, g.y, b.y), act);
        d.z *= act_grad(fmaf(xh.z, g.z, b.z), act); d.w *= act_grad(fmaf(xh.w, g.w, b.w), act);
    }
    return d;
}

template <int MAXV>
__global__ void __launch_bounds__(256, 3) layernorm_bwd_dx_kernel(const void* __restrict__ dy, int dy_dtype, const float* __restrict__ x,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ mean, const float* __restrict__ rstd, int act, float p,
                                                                  uint64_t seed, const unsigned long long* __restrict__ seed_off,
                                                                  const float* __restrict__ addend, float* __restrict__ dx, int M, int D) {
    pdl_enter();
    if (seed_off) seed += *seed_off;
    const int lane = threadIdx.x & 31, row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int nv = D >> 2;
    const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.f;
    const float mu = mean[row], rs = rstd[row];
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
    float4 xh[MAXV], dh[MAXV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < nv) {
            const float4 xv = xr[c], g = reinterpret_cast<const float4*>(gamma)[c];
            xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
            const float4 d = ln_bwd_dterm(load4(dy, dy_dtype, (size_t)row * D + 4 * c), xh[i], g, beta, c, act,
                                          drop4(((size_t)row * D >> 2) + c, p, inv_keep, seed));
            dh[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
            s1 += (dh[i].x + dh[i].y) + (dh[i].z + dh[i].w);
            s2 += (dh[i].x * xh[i].x + dh[i].y * xh[i].y) + (dh[i].z * xh[i].z + dh[i].w * xh[i].w);
        }
    }
    s1 = warp_sum(s1) / (float)D; s2 = warp_sum(s2) / (float)D;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < nv) {
            float4 o;
            o.x = rs * (dh[i].x - s1 - xh[i].x * s2); o.y = rs * (dh[i].y - s1 - xh[i].y * s2);
            o.z = rs * (dh[i].z - s1 - xh[i].z * s2); o.w = rs * (dh[i].w - s1 - xh[i].w * s2);
            if (addend) {          // the gradient that reaches x past this LayerNorm (a pre-LN residual branch): summed here, not in an extra pass
                const float4 a = reinterpret_cast<const float4*>(addend + (size_t)row * D)[c];
                o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
            }
            reinterpret_cast<float4*>(dx + (size_t)row * D)[c] = o;
        }
    }
}

constexpr int LNP_CHUNK_ROWS = 256;
__global__ void __launch_bounds__(256) layernorm_bwd_param_kernel(const void* __restrict__ dy, int dy_dtype, const float* __restrict__ x,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ mean, const float* __restrict__ rstd, int act, float p,
                                                                  uint64_t seed, const unsigned long long* __restrict__ seed_off,
                                                                  float* __restrict__ part, int M, int D) {
    pdl_enter();
    if (seed_off) seed += *seed_off;
    __shared__ float4 red[2][8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;                           // float4 column index
    const int m0 = blockIdx.y * LNP_CHUNK_ROWS, m1 = min(M, m0 + LNP_CHUNK_ROWS);
    const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.f;
    float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag;
    if (4 * c < D) {
        const float4 g = reinterpret_cast<const float4*>(gamma)[c];
#pragma unroll 4
        for (int m = m0 + ty; m < m1; m += 8) {
            const float mu = mean[m], rs = rstd[m];
            const float4 xv = reinterpret_cast<const float4*>(x + (size_t)m * D)[c];
            const float4 xh = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
            const float4 d = ln_bwd_dterm(load4(dy, dy_dtype, (size_t)m * D + 4 * c), xh, g, beta, c, act, drop4(((size_t)m * D >> 2) + c, p, inv_keep, seed));
            ab.x += d.x; ab.y += d.y; ab.z += d.z; ab.w += d.w;
            ag.x = fmaf(d.x, xh.x, ag.x); ag.y = fmaf(d.y, xh.y, ag.y); ag.z = fmaf(d.z, xh.z, ag.z); ag.w = fmaf(d.w, xh.w, ag.w);
        }
    }
    red[0][ty][tx] = ag; red[1][ty][tx] = ab;
    __syncthreads();
    if (ty < 2 && 4 * c < D) {                                    // ty 0: dgamma terms, ty 1: dbeta terms; fixed order over the 8 row lanes
        float4 a = red[ty][0][tx];
#pragma unroll
        for (int w = 1; w < 8; ++w) { const float4 b = red[ty][w][tx]; a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
        reinterpret_cast<float4*>(part + ((size_t)blockIdx.y * 2 + ty) * D)[c] = a;
    }
}
__global__ void __launch_bounds__(256) layernorm_param_reduce_kernel(const float* __restrict__ part, int nparts, int D, float* __restrict__ dgamma,
                                                                     float* __restrict__ dbeta) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * D) return;
    const int which = i / D, c = i - which * D;
    float a = 0.f;
    for (int k = 0; k < nparts; ++k) a += part[((size_t)k * 2 + which) * D + c];
    (which == 0 ? dgamma : dbeta)[c] = a;
}

// ------------------------------------------------------------------------------------------------ activation (+dropout), GLU
__global__ void __launch_bounds__(256) act_fwd_kernel(const float* __restrict__ x, int act, float p, uint64_t seed, const unsigned long long* __restrict__ seed_off,
                                                      float* __restrict__ y32, __nv_bfloat16* __restrict__ y16, size_t n4) {
    pdl_enter();
    if (seed_off) seed += *seed_off;
    const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.f;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (size_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(x)[q], k = drop4(q, p, inv_keep, seed);
        const float4 o = make_float4(act_fwd(v.x, act) * k.x, act_fwd(v.y, act) * k.y, act_fwd(v.z, act) * k.z, act_fwd(v.w, act) * k.w);
        if (y32) reinterpret_cast<float4*>(y32)[q] = o;
        if (y16) store_bf16x4(y16 + 4 * q, o);
    }
}
__global__ void __launch_bounds__(256) act_bwd_kernel(const void* __restrict__ dy, int dy_dtype, const float* __restrict__ x, int act, float p,
                                                      uint64_t seed, const unsigned long long* __restrict__ seed_off, float* __restrict__ dx, size_t n4) {
    pdl_enter();
    if (seed_off) seed += *seed_off;
    const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.f;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (size_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(x)[q], k = drop4(q, p, inv_keep, seed), d = load4(dy, dy_dtype, 4 * q);
        reinterpret_cast<float4*>(dx)[q] = make_float4(d.x * k.x * act_grad(v.x, act), d.y * k.y * act_grad(v.y, act),
                                                       d.z * k.z * act_grad(v.z, act), d.w * k.w * act_grad(v.w, act));
    }
}
__device__ __forceinline__ float sigm(float v) { return 1.0f / (1.0f + __expf(-v)); }
__global__ void __launch_bounds__(256) glu_fwd_kernel(const float* __restrict__ u, float* __restrict__ g, int M, int D) {
    pdl_enter();
    const size_t n4 = (size_t)M * D >> 2;
    const int d4 = D >> 2;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (size_t)gridDim.x * blockDim.x) {
        const size_t row = q / d4;
        const int c = (int)(q - row * d4);
        const float4 a = reinterpret_cast<const float4*>(u + row * 2 * D)[c], b = reinterpret_cast<const float4*>(u + row * 2 * D + D)[c];
        reinterpret_cast<float4*>(g)[q] = make_float4(a.x * sigm(b.x), a.y * sigm(b.y), a.z * sigm(b.z), a.w * sigm(b.w));
    }
}
__global__ void __launch_bounds__(256) glu_bwd_kernel(const float* __restrict__ dg, const float* __restrict__ u, float* __restrict__ du, int M, int D) {
    pdl_enter();
    const size_t n4 = (size_t)M * D >> 2;
    const int d4 = D >> 2;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (size_t)gridDim.x * blockDim.x) {
        const size_t row = q / d4;
        const int c = (int)(q - row * d4);
        const float4 a = reinterpret_cast<const float4*>(u + row * 2 * D)[c], b = reinterpret_cast<const float4*>(u + row * 2 * D + D)[c];
        const float4 d = reinterpret_cast<const float4*>(dg)[q];
        const float4 s = make_float4(sigm(b.x), sigm(b.y), sigm(b.z), sigm(b.w));
        reinterpret_cast<float4*>(du + row * 2 * D)[c] = make_float4(d.x * s.x, d.y * s.y, d.z * s.z, d.w * s.w);
        reinterpret_cast<float4*>(du + row * 2 * D + D)[c] = make_float4(d.x * a.x * s.x * (1.f - s.x), d.y * a.y * s.y * (1.f - s.y),
                                                                         d.z * a.z * s.z * (1.f - s.z), d.w * a.w * s.w * (1.f - s.w));
    }
}

// ------------------------------------------------------------------------------------------------ residual: out = x + scale * path[b] * drop(y)
__global__ void __launch_bounds__(256) residual_kernel(const float* __restrict__ x, const float* __restrict__ y, float scale, float p, uint64_t seed,
                                                       float p_path, uint64_t path_seed, const unsigned long long* __restrict__ seed_off, size_t elems_per_sample,
                                                       float* __restrict__ out, size_t n4) {
    pdl_enter();
    if (seed_off) { seed += *seed_off; path_seed += *seed_off; }
    const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.f;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (size_t)gridDim.x * blockDim.x) {
        const float f = scale * path_factor((int)((4 * q) / elems_per_sample), p_path, path_seed);
        const float4 v = reinterpret_cast<const float4*>(y)[q], k = drop4(q, p, inv_keep, seed);
        float4 o = make_float4(v.x * k.x * f, v.y * k.y * f, v.z * k.z * f, v.w * k.w * f);
        if (x) { const float4 r = reinterpret_cast<const float4*>(x)[q]; o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
        reinterpret_cast<float4*>(out)[q] = o;
    }
}

// ------------------------------------------------------------------------------------------------ depthwise convolution over time
// y[b,t,d] = bias[d] + sum_j w[d][j] x[b, t + j - k/2, d]   (flip: w[d][k-1-j], the data gradient).  A thread owns one channel
// and DW_TT consecutive frames: the input window lives in registers, the taps in shared memory (tap-major: conflict-free).
constexpr int DW_TT = 16, DW_MAXK = 32;
__global__ void __launch_bounds__(128) dwconv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                         float* __restrict__ y, int T, int D, int k, int flip, int w_shared) {
    pdl_enter();
    extern __shared__ float ws[];                       // [128][k] as stored (odd k: a thread's taps sit 'k' words apart -> conflict-free)
    const int d = blockIdx.x * 128 + threadIdx.x, b = blockIdx.z, t0 = blockIdx.y * DW_TT, pad = k / 2;
    float win[DW_TT + DW_MAXK - 1], wreg[DW_MAXK];
    const float* xb = x + (size_t)b * T * D + min(d, D - 1);
#pragma unroll
    for (int i = 0; i < DW_TT + DW_MAXK - 1; ++i) {     // the input window is requested first: it arrives while the taps are staged
        const int t = t0 + i - pad;
        win[i] = (i < DW_TT + k - 1 && t >= 0 && t < T) ? xb[(size_t)t * D] : 0.f;
    }
    {
        const int nd = min(128, D - blockIdx.x * 128);
        const float* wsrc = w + (w_shared ? 0 : (size_t)blockIdx.x * 128 * k);
        for (int i = threadIdx.x; i < nd * k; i += 128) ws[i] = w_shared ? wsrc[i % k] : wsrc[i];      // coalesced
    }
    __syncthreads();
    if (d >= D) return;
#pragma unroll
    for (int j = 0; j < DW_MAXK; ++j) wreg[j] = j < k ? ws[threadIdx.x * k + (flip ? k - 1 - j : j)] : 0.f;   // taps beyond k are zero: branch-free inner loop
    const float bv = bias ? bias[d] : 0.f;
#pragma unroll
    for (int tt = 0; tt < DW_TT; ++tt) {
        float acc = bv;
#pragma unroll
        for (int j = 0; j < DW_MAXK; ++j) acc = fmaf(wreg[j], win[tt + j], acc);
        if (t0 + tt < T) y[((size_t)b * T + t0 + tt) * D + d] = acc;
    }
}
// dw[d][j] = sum_{b,t} dy[b,t,d] x[b, t + j - k/2, d], db[d] = sum dy: CTA = (128 channels, a chunk of utterances); per-chunk partials
// reduced in fixed order by dwconv_w_reduce_kernel.
__global__ void __launch_bounds__(128) dwconv_bwd_w_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ part, int B,
                                                           int T, int D, int k, int b_per_cta) {
    pdl_enter();
    // thread = one channel; per 16-frame tile the gradient rows and the input window live in registers and feed all k tap sums
    const int d = blockIdx.x * 128 + threadIdx.x, pad = k / 2;
    if (d >= D) return;
    float acc[DW_MAXK], accb = 0.f;
#pragma unroll
    for (int j = 0; j < DW_MAXK; ++j) acc[j] = 0.f;
    const int b0 = blockIdx.y * b_per_cta, b1 = min(B, b0 + b_per_cta);
    for (int b = b0; b < b1; ++b) {
        const float* dyb = dy + (size_t)b * T * D + d;
        const float* xb = x + (size_t)b * T * D + d;
        for (int t0 = 0; t0 < T; t0 += DW_TT) {
            float g[DW_TT], win[DW_TT + DW_MAXK - 1];
#pragma unroll
            for (int i = 0; i < DW_TT; ++i) g[i] = t0 + i < T ? dyb[(size_t)(t0 + i) * D] : 0.f;
#pragma unroll
            for (int i = 0; i < DW_TT + DW_MAXK - 1; ++i) {
                const int t = t0 + i - pad;
                win[i] = (i < DW_TT + k - 1 && t >= 0 && t < T) ? xb[(size_t)t * D] : 0.f;
            }
#pragma unroll
            for (int i = 0; i < DW_TT; ++i) {
                accb += g[i];
#pragma unroll
                for (int j = 0; j < DW_MAXK; ++j) acc[j] = fmaf(g[i], win[i + j], acc[j]);      // taps >= k see zeros of the window
            }
        }
    }
    float* pr = part + (size_t)blockIdx.y * D * (k + 1) + (size_t)d * (k + 1);
#pragma unroll
    for (int j = 0; j < DW_MAXK; ++j)
        if (j < k) pr[j] = acc[j];
    pr[k] = accb;
}
__global__ void __launch_bounds__(256) dwconv_w_reduce_kernel(const float* __restrict__ part, int nparts, int D, int k, float* __restrict__ dw,
                                                              float* __restrict__ db) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D * (k + 1)) return;
    float a = 0.f;
    for (int p = 0; p < nparts; ++p) a += part[(size_t)p * D * (k + 1) + i];
    const int d = i / (k + 1), j = i - d * (k + 1);
    if (j < k) dw[(size_t)d * k + j] = a;
    else if (db) db[d] = a;
}

// strided depthwise conv without padding: y[b,j,c] = sum_k w[c][k] x[b, j*S + k, c], j < T' = (T-K)/S + 1
__global__ void __launch_bounds__(128) strided_dwconv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y32,
                                                                 __nv_bfloat16* __restrict__ y16, int T, int N, int K, int S, int Tp) {
    pdl_enter();
    const int c = blockIdx.x * 128 + threadIdx.x, j = blockIdx.y, b = blockIdx.z;
    if (c >= N) return;
    const float* xb = x + ((size_t)b * T + (size_t)j * S) * N + c;
    const float* wc = w + (size_t)c * K;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(wc[k], xb[(size_t)k * N], acc);
    const size_t o = ((size_t)b * Tp + j) * N + c;
    if (y32) y32[o] = acc;
    if (y16) y16[o] = __float2bfloat16_rn(acc);
}
// dx[b,t,c] = sum_{(j,k): jS+k=t} dy[b,j,c] w[c][k]   (gather form: deterministic)
__global__ void __launch_bounds__(128) strided_dwconv_bwd_x_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                                                                   int T, int N, int K, int S, int Tp) {
    pdl_enter();
    const int c = blockIdx.x * 128 + threadIdx.x, t = blockIdx.y, b = blockIdx.z;
    if (c >= N) return;
    float acc = 0.f;
    const int jhi = min(Tp - 1, t / S);
    for (int j = jhi; j >= 0 && t - j * S < K; --j) acc = fmaf(dy[((size_t)b * Tp + j) * N + c], w[(size_t)c * K + t - j * S], acc);
    dx[((size_t)b * T + t) * N + c] = acc;
}
// dw[c][k] partials per utterance chunk: thread = (channel, tap)
__global__ void __launch_bounds__(256) strided_dwconv_bwd_w_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ part,
                                                                   int B, int T, int N, int K, int S, int Tp, int b_per_cta) {
    pdl_enter();
    const int lane = threadIdx.x & 31, c = blockIdx.x * 32 + lane;
    if (c >= N) return;
    const int b0 = blockIdx.y * b_per_cta, b1 = min(B, b0 + b_per_cta);
    for (int k = threadIdx.x >> 5; k < K; k += 8) {
        float acc = 0.f;
        for (int b = b0; b < b1; ++b)
            for (int j = 0; j < Tp; ++j) acc = fmaf(dy[((size_t)b * Tp + j) * N + c], x[((size_t)b * T + (size_t)j * S + k) * N + c], acc);
        part[(size_t)blockIdx.y * N * (K + 1) + (size_t)c * (K + 1) + k] = acc;
    }
    if ((threadIdx.x >> 5) == 0) part[(size_t)blockIdx.y * N * (K + 1) + (size_t)c * (K + 1) + K] = 0.f;
}

// ------------------------------------------------------------------------------------------------ SpecAugment bands + positional encoding
// out[b,t,d] = (masked ? 0 : z[b,t,d]) + pe[t,d] (pe == NULL: the backward, out = masked ? 0 : z); bands[8] = {f0,f1, f0,f1, t0,t1, t0,t1}
struct Bands { int v[8]; };
__global__ void __launch_bounds__(256) posenc_mask_kernel(const float* __restrict__ z, const float* __restrict__ pe, Bands bands,
                                                          const int* __restrict__ bands_dev, float* __restrict__ out, int T, int D, size_t n4) {
    pdl_enter();
    if (bands_dev) {
#pragma unroll
        for (int i = 0; i < 8; ++i) bands.v[i] = bands_dev[i];
    }
    const int d4 = D >> 2;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (size_t)gridDim.x * blockDim.x) {
        const size_t row = q / d4;
        const int c = (int)(q - row * d4) * 4, t = (int)(row % T);
        float4 v = reinterpret_cast<const float4*>(z)[q];
        const bool tm = (t >= bands.v[4] && t < bands.v[5]) || (t >= bands.v[6] && t < bands.v[7]);
        float* vv = reinterpret_cast<float*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = c + i;
            if (tm || (f >= bands.v[0] && f < bands.v[1]) || (f >= bands.v[2] && f < bands.v[3])) vv[i] = 0.f;
        }
        if (pe) { const float4 e = reinterpret_cast<const float4*>(pe + (size_t)t * D)[c >> 2]; v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w; }
        reinterpret_cast<float4*>(out)[q] = v;
    }
}

// ------------------------------------------------------------------------------------------------ small utilities
// out[d] = sum over the rows b with index[b] == d, in row order (dayWeights / dayBias gradients: index_select backward)
__global__ void __launch_bounds__(256) index_reduce_kernel(const float* __restrict__ part, const int64_t* __restrict__ index, int B, size_t n, int n_out,
                                                           float* __restrict__ out) {
    pdl_enter();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int d = blockIdx.y;
    if (i >= n) return;
    float a = 0.f;
    for (int b = 0; b < B; ++b)
        if (index[b] == d) a += part[(size_t)b * n + i];
    out[(size_t)d * n + i] = a;
}
__global__ void __launch_bounds__(256) axpb_kernel(const float* __restrict__ x, float a, float b, float* __restrict__ y, size_t n) {
    pdl_enter();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = fmaf(a, x[i], b);
}
// out = a * in + b*out0 ... single-CTA deterministic sum (the KL term of the label-smoothed loss is a plain sum of log-probs)
__global__ void __launch_bounds__(1024) sum_kernel(const float* __restrict__ x, size_t n, float scale, float add, int accumulate, float* __restrict__ out) {
    pdl_enter();
    __shared__ float sm[32];
    float a = 0.f;
    for (size_t i = threadIdx.x; i < n; i += 1024) a += x[i];
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        a = warp_sum(sm[threadIdx.x]);
        if (threadIdx.x == 0) *out = (accumulate ? *out : 0.f) + fmaf(a, scale, add);
    }
}
// dlogits = dlp - softmax * sum_c dlp   (log_softmax backward, rows of C)
__global__ void __launch_bounds__(256) log_softmax_bwd_kernel(const float* __restrict__ lp, const float* __restrict__ dlp, float* __restrict__ dl,
                                                              int64_t rows, int C) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += dlp[row * C + c];
    s = warp_sum(s);
    for (int c = lane; c < C; c += 32) dl[row * C + c] = dlp[row * C + c] - __expf(lp[row * C + c]) * s;
}


// bf16 copy of a f32 [M,N] gradient (the tcgen05 operand of the dgrad / wgrad GEMMs) AND its column sums (the bias gradient) from ONE
// read: CTA = 128 columns x a chunk of rows, thread = 4 columns x every 8th row; per-chunk partial sums, reduced in fixed order.
constexpr int CC_CHUNK_ROWS = 256;
__global__ void __launch_bounds__(256) cast_colsum_kernel(const float* __restrict__ src, int M, int N, __nv_bfloat16* __restrict__ dst, int ld_dst,
                                                          float* __restrict__ part) {
    pdl_enter();
    __shared__ float4 red[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.x * 128 + 4 * tx;
    const int m0 = blockIdx.y * CC_CHUNK_ROWS, m1 = min(M, m0 + CC_CHUNK_ROWS);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) {
#pragma unroll 4
        for (int m = m0 + ty; m < m1; m += 8) {
            const float4 v = *reinterpret_cast<const float4*>(src + (size_t)m * N + n);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            store_bf16x4(dst + (size_t)m * ld_dst + n, v);
        }
    }
    red[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && n < N) {
#pragma unroll
        for (int w = 1; w < 8; ++w) { const float4 b = red[w][tx]; a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
        *reinterpret_cast<float4*>(part + (size_t)blockIdx.y * N + n) = a;
    }
}
__global__ void __launch_bounds__(256) colsum_chunks_kernel(const float* __restrict__ part, int N, int chunks, float* __restrict__ out) {
    pdl_enter();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += part[(size_t)c * N + n];
    out[n] = s;
}

static inline int ew_blocks(size_t n4) { return (int)std::min<size_t>(std::max<size_t>(cdivz(n4, 256), 1), (size_t)sm_count() * 16); }

}  // namespace nsd

extern "C" {

using namespace nsd;

int nsd_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int act, float p_drop, uint64_t seed, float* y_f32,
                      void* y_bf16, float* mean, float* rstd, int M, int D, void* stream) {
    NSD_CHECK_ARG(M >= 0 && D > 0 && D % 4 == 0 && D <= 128 * LN_MAXV, "layernorm_fwd: bad sizes M=%d D=%d (D %% 4 == 0, D <= %d)", M, D, 128 * LN_MAXV);
    NSD_CHECK_ARG(x && gamma && beta && mean && rstd && (y_f32 || y_bf16) && act >= 0 && act <= 3 && p_drop >= 0.f && p_drop < 1.f, "layernorm_fwd: bad argument");
    if (M == 0) return NSD_OK;
#define NSD_LN_FWD(V) nsd::launch_k(layernorm_fwd_kernel<V>, cdiv(M, 8), 256, 0, (cudaStream_t)stream, x, gamma, beta, eps, act, p_drop, seed, seed_offset_ptr(), y_f32, (__nv_bfloat16*)y_bf16, mean, rstd, M, D)
    if (D <= 512) NSD_LN_FWD(4); else if (D <= 1024) NSD_LN_FWD(8); else NSD_LN_FWD(16);
#undef NSD_LN_FWD
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}
size_t nsd_layernorm_bwd_workspace(int M, int D) { return sizeof(float) * 2 * (size_t)D * (size_t)cdiv(std::max(M, 1), LNP_CHUNK_ROWS); }
int nsd_layernorm_bwd(const void* dy, int dy_dtype, const float* x, const float* gamma, const float* beta, const float* mean, const float* rstd, int act,
                      float p_drop, uint64_t seed, const float* dx_addend, float* dx, float* dgamma, float* dbeta, int M, int D, void* workspace,
                      size_t workspace_bytes, void* stream) {
    NSD_CHECK_ARG(M >= 0 && D > 0 && D % 4 == 0 && D <= 128 * LN_MAXV, "layernorm_bwd: bad sizes M=%d D=%d", M, D);
    NSD_CHECK_ARG(dy && x && gamma && beta && mean && rstd && dx && dgamma && dbeta && (dy_dtype == NSD_F32 || dy_dtype == NSD_BF16), "layernorm_bwd: bad argument");
    if (workspace_bytes < nsd_layernorm_bwd_workspace(M, D) || !workspace) { set_error("layernorm_bwd: workspace too small"); return NSD_ERR_WORKSPACE; }
    if (M == 0) return NSD_OK;
    const int parts = cdiv(M, LNP_CHUNK_ROWS);
    cudaStream_t st = (cudaStream_t)stream;
#define NSD_LN_BWD(V) nsd::launch_k(layernorm_bwd_dx_kernel<V>, cdiv(M, 8), 256, 0, st, dy, dy_dtype, x, gamma, beta, mean, rstd, act, p_drop, seed, seed_offset_ptr(), dx_addend, dx, M, D)
    if (D <= 512) NSD_LN_BWD(4); else if (D <= 1024) NSD_LN_BWD(8); else NSD_LN_BWD(16);
#undef NSD_LN_BWD
    NSD_LAUNCH_CHECK();
    nsd::launch_k(layernorm_bwd_param_kernel, dim3(cdiv(D, 128), parts), 256, 0, st, dy, dy_dtype, x, gamma, beta, mean, rstd, act, p_drop, seed, seed_offset_ptr(),
                                                                          (float*)workspace, M, D);
    NSD_LAUNCH_CHECK();
    nsd::launch_k(layernorm_param_reduce_kernel, cdiv(2 * D, 256), 256, 0, (cudaStream_t)stream, (const float*)workspace, parts, D, dgamma, dbeta);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}

int nsd_act_fwd(const float* x, int act, float p_drop, uint64_t seed, float* y_f32, void* y_bf16, size_t n, void* stream) {
    NSD_CHECK_ARG(x && (y_f32 || y_bf16) && n % 4 == 0 && act >= 0 && act <= 3 && p_drop >= 0.f && p_drop < 1.f, "act_fwd: bad argument (n %% 4 == 0)");
    if (n == 0) return NSD_OK;
    nsd::launch_k(act_fwd_kernel, ew_blocks(n / 4), 256, 0, (cudaStream_t)stream, x, act, p_drop, seed, seed_offset_ptr(), y_f32, (__nv_bfloat16*)y_bf16, n / 4);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}
int nsd_act_bwd(const void* dy, int dy_dtype, const float* x, int act, float p_drop, uint64_t seed, float* dx, size_t n, void* stream) {
    NSD_CHECK_ARG(dy && x && dx && n % 4 == 0 && act >= 0 && act <= 3 && (dy_dtype == NSD_F32 || dy_dtype == NSD_BF16), "act_bwd: bad argument");
    if (n == 0) return NSD_OK;
    nsd::launch_k(act_bwd_kernel, ew_blocks(n / 4), 256, 0, (cudaStream_t)stream, dy, dy_dtype, x, act, p_drop, seed, seed_offset_ptr(), dx, n / 4);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}
int nsd_glu_fwd(const float* u, float* g, int M, int D, void* stream) {
    NSD_CHECK_ARG(u && g && M >= 0 && D > 0 && D % 4 == 0, "glu_fwd: bad argument");
    if (M == 0) return NSD_OK;
    nsd::launch_k(glu_fwd_kernel, ew_blocks((size_t)M * D / 4), 256, 0, (cudaStream_t)stream, u, g, M, D);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}
int nsd_glu_bwd(const float* dg, const float* u, float* du, int M, int D, void* stream) {
    NSD_CHECK_ARG(dg && u && du && M >= 0 && D > 0 && D % 4 == 0, "glu_bwd: bad argument");
    if (M == 0) return NSD_OK;
    nsd::launch_k(glu_bwd_kernel, ew_blocks((size_t)M * D / 4), 256, 0, (cudaStream_t)stream, dg, u, du, M, D);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}
int nsd_residual(const float* x, const float* y, float scale, float p_drop, uint64_t seed, float p_path, uint64_t path_seed, int64_t elems_per_sample,
                 float* out, size_t n, void* stream) {
    NSD_CHECK_ARG(y && out && n % 4 == 0 && elems_per_sample > 0 && elems_per_sample % 4 == 0 && p_drop >= 0.f && p_drop < 1.f && p_path >= 0.f && p_path < 1.f,
                  "residual: bad argument");
    if (n == 0) return NSD_OK;
    nsd::launch_k(residual_kernel, ew_blocks(n / 4), 256, 0, (cudaStream_t)stream, x, y, scale, p_drop, seed, p_path, path_seed, seed_offset_ptr(), (size_t)elems_per_sample, out, n / 4);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}

int nsd_dwconv_fwd(const float* x, const float* w, const float* bias, float* y, int B, int T, int D, int k, int flip, int w_shared, void* stream) {
    NSD_CHECK_ARG(x && w && y && B >= 0 && T > 0 && D > 0 && k >= 1 && k <= DW_MAXK && (k & 1), "dwconv_fwd: bad argument (odd k <= %d)", DW_MAXK);
    if (B == 0) return NSD_OK;
    nsd::launch_k(dwconv_fwd_kernel, dim3(cdiv(D, 128), cdiv(T, DW_TT), B), 128, sizeof(float) * k * 128, (cudaStream_t)stream, x, w, bias, y, T, D, k, flip, w_shared);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}
static int dw_parts(int B) { return std::min(B, 64); }
size_t nsd_dwconv_bwd_w_workspace(int B, int D, int k) { return sizeof(float) * (size_t)dw_parts(std::max(B, 1)) * D * (k + 1); }
int nsd_dwconv_bwd_w(const float* dy, const float* x, float* dw, float* db, int B, int T, int D, int k, void* workspace, size_t workspace_bytes,
                     void* stream) {
    NSD_CHECK_ARG(dy && x && dw && B >= 1 && T > 0 && D > 0 && k >= 1 && k <= DW_MAXK && (k & 1), "dwconv_bwd_w: bad argument");
    if (!workspace || workspace_bytes < nsd_dwconv_bwd_w_workspace(B, D, k)) { set_error("dwconv_bwd_w: workspace too small"); return NSD_ERR_WORKSPACE; }
    const int parts = dw_parts(B), per = cdiv(B, parts);
    nsd::launch_k(dwconv_bwd_w_kernel, dim3(cdiv(D, 128), cdiv(B, per)), 128, 0, (cudaStream_t)stream, dy, x, (float*)workspace, B, T, D, k, per);
    NSD_LAUNCH_CHECK();
    nsd::launch_k(dwconv_w_reduce_kernel, cdiv(D * (k + 1), 256), 256, 0, (cudaStream_t)stream, (const float*)workspace, cdiv(B, per), D, k, dw, db);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}
int nsd_strided_dwconv_fwd(const float* x, const float* w, float* y_f32, void* y_bf16, int B, int T, int N, int K, int S, void* stream) {
    NSD_CHECK_ARG(x && w && (y_f32 || y_bf16) && B >= 0 && N > 0 && K >= 1 && S >= 1 && T >= K, "strided_dwconv_fwd: bad argument (T >= K)");
    if (B == 0) return NSD_OK;
    const int Tp = (T - K) / S + 1;
    nsd::launch_k(strided_dwconv_fwd_kernel, dim3(cdiv(N, 128), Tp, B), 128, 0, (cudaStream_t)stream, x, w, y_f32, (__nv_bfloat16*)y_bf16, T, N, K, S, Tp);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}
size_t nsd_strided_dwconv_bwd_workspace(int B, int N, int K) { return sizeof(float) * (size_t)dw_parts(std::max(B, 1)) * N * (K + 1); }
int nsd_strided_dwconv_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, int B, int T, int N, int K, int S, void* workspace,
                           size_t workspace_bytes, void* stream) {
    NSD_CHECK_ARG(dy && x && w && dx && dw && B >= 1 && N > 0 && K >= 1 && S >= 1 && T >= K, "strided_dwconv_bwd: bad argument");
    if (!workspace || workspace_bytes < nsd_strided_dwconv_bwd_workspace(B, N, K)) { set_error("strided_dwconv_bwd: workspace too small"); return NSD_ERR_WORKSPACE; }
    const int Tp = (T - K) / S + 1, parts = dw_parts(B), per = cdiv(B, parts);
    nsd::launch_k(strided_dwconv_bwd_x_kernel, dim3(cdiv(N, 128), T, B), 128, 0, (cudaStream_t)stream, dy, w, dx, T, N, K, S, Tp);
    NSD_LAUNCH_CHECK();
    nsd::launch_k(strided_dwconv_bwd_w_kernel, dim3(cdiv(N, 32), cdiv(B, per)), 256, 0, (cudaStream_t)stream, dy, x, (float*)workspace, B, T, N, K, S, Tp, per);
    NSD_LAUNCH_CHECK();
    nsd::launch_k(dwconv_w_reduce_kernel, cdiv(N * (K + 1), 256), 256, 0, (cudaStream_t)stream, (const float*)workspace, cdiv(B, per), N, K, dw, nullptr);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}
size_t nsd_cast_colsum_workspace(int M, int N) { return sizeof(float) * (size_t)cdiv(std::max(M, 1), CC_CHUNK_ROWS) * (size_t)std::max(N, 1); }
int nsd_cast_colsum(const float* src, int M, int N, void* dst_bf16, int ld_dst, float* colsum, void* workspace, size_t workspace_bytes, void* stream) {
    NSD_CHECK_ARG(src && dst_bf16 && colsum && M >= 1 && N >= 4 && N % 4 == 0 && ld_dst >= N && ld_dst % 4 == 0, "cast_colsum: bad argument (N %% 4 == 0)");
    if (!workspace || workspace_bytes < nsd_cast_colsum_workspace(M, N)) { set_error("cast_colsum: workspace too small"); return NSD_ERR_WORKSPACE; }
    const int chunks = cdiv(M, CC_CHUNK_ROWS);
    nsd::launch_k(cast_colsum_kernel, dim3(cdiv(N, 128), chunks), 256, 0, (cudaStream_t)stream, src, M, N, (__nv_bfloat16*)dst_bf16, ld_dst, (float*)workspace);
    NSD_LAUNCH_CHECK();
    nsd::launch_k(colsum_chunks_kernel, cdiv(N, 256), 256, 0, (cudaStream_t)stream, (const float*)workspace, N, chunks, colsum);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}
int nsd_posenc_mask(const float* z, const float* pe, const int* bands8, const int* bands8_dev, float* out, int B, int T, int D, void* stream) {
    NSD_CHECK_ARG(z && out && (bands8 || bands8_dev) && B >= 0 && T > 0 && D > 0 && D % 4 == 0, "posenc_mask: bad argument");
    if (B == 0) return NSD_OK;
    Bands bd;
    for (int i = 0; i < 8; ++i) bd.v[i] = bands8 ? bands8[i] : 0;
    const size_t n4 = (size_t)B * T * D / 4;
    nsd::launch_k(posenc_mask_kernel, ew_blocks(n4), 256, 0, (cudaStream_t)stream, z, pe, bd, bands8_dev, out, T, D, n4);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}
int nsd_index_reduce(const float* partial, const int64_t* index, int B, size_t n, int n_out, float* out, void* stream) {
    NSD_CHECK_ARG(partial && index && out && B >= 0 && n_out >= 1, "index_reduce: bad argument");
    if (n == 0) return NSD_OK;
    nsd::launch_k(index_reduce_kernel, dim3((unsigned)cdivz(n, 256), n_out), 256, 0, (cudaStream_t)stream, partial, index, B, n, n_out, out);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}
int nsd_axpb(const float* x, float a, float b, float* y, size_t n, void* stream) {
    NSD_CHECK_ARG(x && y, "axpb: null pointer");
    if (n == 0) return NSD_OK;
    nsd::launch_k(axpb_kernel, ew_blocks(n), 256, 0, (cudaStream_t)stream, x, a, b, y, n);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}
int nsd_sum_f32(const float* x, size_t n, float scale, float add, int accumulate, float* out, void* stream) {
    NSD_CHECK_ARG(x && out, "sum_f32: null pointer");
    nsd::launch_k(sum_kernel, 1, 1024, 0, (cudaStream_t)stream, x, n, scale, add, accumulate, out);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}
int nsd_log_softmax_bwd(const float* lp, const float* dlp, float* dlogits, int64_t rows, int C, void* stream) {
    NSD_CHECK_ARG(lp && dlp && dlogits && rows >= 0 && C > 0, "log_softmax_bwd: bad argument");
    if (rows == 0) return NSD_OK;
    nsd::launch_k(log_softmax_bwd_kernel, (unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream, lp, dlp, dlogits, rows, C);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}

}  // extern "C"
