// Multi-head self-attention of the Conformer path (reference transformer_ctc.py:215-217, 248-251: nn.MultiheadAttention with
// a boolean key-padding mask and dropout on the attention weights) as strided batched GEMMs + a masked row softmax, and the
// AdamW / gradient-norm pieces of its training step (neural_decoder_trainer.py:144-162, 255-259).
//
// At the competition shape a head is a 118 x 118 x 128 problem and there are B x heads = 512 of them per layer: far below a
// tcgen05 tile pipeline's efficient size and < 2 % of the step's FLOPs (the 1024-wide projections around it run on the
// tcgen05 GEMM).  nsd_bgemm therefore uses warp-level mma.sync (bf16 operands, fp32 accumulate) on 64 x 64 tiles, one CTA per
// (tile, batch entry), with operands addressed by general (row, column, batch0, batch1) strides so that Q, K, V are read in
// place from the packed projection output [B*T, 3D] and the result lands in place in [B*T, D] -- no head split / merge
// copies.  fp32 operands (the parity mode) take an FFMA path in the same kernel.
#include <algorithm>
#include <math.h>
#include <type_traits>

#include "common.cuh"

namespace nsd {

struct BgemmParams {
    const void* A; const void* B; void* C; const float* bias; const int64_t* b_index;
    long long a_rs, a_cs, a_b0, a_b1, b_rs, b_cs, b_b0, b_b1, c_rs, c_b0, c_b1, bias_b0;
    int a_dtype, b_dtype, c_dtype, M, N, K, nb1;
    float alpha;
};

__device__ __forceinline__ float ld_elem(const void* p, int dtype, long long i) {
    return dtype == NSD_F32 ? reinterpret_cast<const float*>(p)[i] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}

constexpr int BG_T = 64, BG_K = 32, BG_PAD = 8;

// loads a [64 rows x 32 k] tile of an operand addressed as base[row*rs + k*cs] into smem[row][k] (zero outside), threads
// walking the contiguous direction of the source
template <typename ST>
__device__ __forceinline__ void bg_load_tile(ST (*sm)[BG_K + BG_PAD], const void* base, int dtype, long long rs, long long cs, int row0, int nrows, int k0,
                                             int K, int tid) {
    if (cs == 1 || rs != 1) {                 // k contiguous (or no contiguous direction): consecutive threads along k
        for (int i = tid; i < BG_T * BG_K; i += 128) {
            const int r = i >> 5, k = i & 31;
            float v = 0.f;
            if (row0 + r < nrows && k0 + k < K) v = ld_elem(base, dtype, (long long)(row0 + r) * rs + (long long)(k0 + k) * cs);
            if constexpr (sizeof(ST) == 2) sm[r][k] = __float2bfloat16_rn(v); else sm[r][k] = v;
        }
    } else {                                  // rows contiguous: consecutive threads along the row index
        for (int i = tid; i < BG_T * BG_K; i += 128) {
            const int k = i >> 6, r = i & 63;
            float v = 0.f;
            if (row0 + r < nrows && k0 + k < K) v = ld_elem(base, dtype, (long long)(row0 + r) + (long long)(k0 + k) * cs);
            if constexpr (sizeof(ST) == 2) sm[r][k] = __float2bfloat16_rn(v); else sm[r][k] = v;
        }
    }
}

template <bool TC>
__global__ void __launch_bounds__(128) bgemm_kernel(const BgemmParams p) {
    using ST = typename std::conditional<TC, __nv_bfloat16, float>::type;
    __shared__ __align__(16) ST As[BG_T][BG_K + BG_PAD];
    __shared__ __align__(16) ST Bs[BG_T][BG_K + BG_PAD];      // [n][k]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int z = blockIdx.z, z0 = z / p.nb1, z1 = z - z0 * p.nb1;
    const int m0 = blockIdx.y * BG_T, n0 = blockIdx.x * BG_T;
    const long long bsel = p.b_index ? p.b_index[z0] : z0;
    const size_t ea = p.a_dtype == NSD_F32 ? 4 : 2, eb = p.b_dtype == NSD_F32 ? 4 : 2;
    const char* A = reinterpret_cast<const char*>(p.A) + (size_t)(z0 * p.a_b0 + z1 * p.a_b1) * ea;
    const char* Bm = reinterpret_cast<const char*>(p.B) + (size_t)(bsel * p.b_b0 + z1 * p.b_b1) * eb;
    const float* bias = p.bias ? p.bias + bsel * p.bias_b0 : nullptr;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;           // warp tile 32 x 32
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;
    for (int k0 = 0; k0 < p.K; k0 += BG_K) {
        bg_load_tile<ST>(As, A, p.a_dtype, p.a_rs, p.a_cs, m0, p.M, k0, p.K, tid);
        bg_load_tile<ST>(Bs, Bm, p.b_dtype, p.b_cs, p.b_rs, n0, p.N, k0, p.K, tid);      // row of Bs = n: (rs, cs) seen from n are (b_cs, b_rs)
        __syncthreads();
        if constexpr (TC) {
#pragma unroll
            for (int ks = 0; ks < BG_K; ks += 16) {
                uint32_t a[2][4], b[2][4];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(&As[wm + 16 * i + (lane & 15)][ks + 8 * (lane >> 4)]);
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a[i][0]), "=r"(a[i][1]), "=r"(a[i][2]), "=r"(a[i][3]) : "r"(addr));
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {     // two n8 tiles per ldmatrix.x4: matrices (n 0-7,k 0-7), (n 0-7,k 8-15), (n 8-15,k 0-7), (n 8-15,k 8-15)
                    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(&Bs[wn + 16 * j + (lane & 7) + 8 * (lane >> 4)][ks + 8 * ((lane >> 3) & 1)]);
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(b[j][0]), "=r"(b[j][1]), "=r"(b[j][2]), "=r"(b[j][3]) : "r"(addr));
                }
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t b0 = b[j >> 1][2 * (j & 1)], b1 = b[j >> 1][2 * (j & 1) + 1];
                        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                     : "+f"(acc[i][j][0]), "+f"(acc[i][j][1]), "+f"(acc[i][j][2]), "+f"(acc[i][j][3])
                                     : "r"(a[i][0]), "r"(a[i][1]), "r"(a[i][2]), "r"(a[i][3]), "r"(b0), "r"(b1));
                    }
            }
        } else {
            // same accumulator ownership as the mma fragments: rows wm + 16 i + g (+8), columns wn + 8 j + 2 c (+1)
            const int g = lane >> 2, c = lane & 3;
            for (int k = 0; k < BG_K; ++k) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float a0 = As[wm + 16 * i + g][k], a1 = As[wm + 16 * i + g + 8][k];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float b0 = Bs[wn + 8 * j + 2 * c][k], b1 = Bs[wn + 8 * j + 2 * c + 1][k];
                        acc[i][j][0] = fmaf(a0, b0, acc[i][j][0]); acc[i][j][1] = fmaf(a0, b1, acc[i][j][1]);
                        acc[i][j][2] = fmaf(a1, b0, acc[i][j][2]); acc[i][j][3] = fmaf(a1, b1, acc[i][j][3]);
                    }
                }
            }
        }
        __syncthreads();
    }
    const int g = lane >> 2, c = lane & 3;
    char* C = reinterpret_cast<char*>(p.C);
    const long long cbase = (long long)z0 * p.c_b0 + (long long)z1 * p.c_b1;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int m = m0 + wm + 16 * i + g + 8 * (e >> 1), n = n0 + wn + 8 * j + 2 * c + (e & 1);
                if (m < p.M && n < p.N) {
                    const float v = fmaf(p.alpha, acc[i][j][e], bias ? bias[n] : 0.f);
                    const long long o = cbase + (long long)m * p.c_rs + n;
                    if (p.c_dtype == NSD_F32) reinterpret_cast<float*>(C)[o] = v;
                    else reinterpret_cast<__nv_bfloat16*>(C)[o] = __float2bfloat16_rn(v);
                }
            }
}

// ---- masked row softmax of the attention scores (in place), optional dropped copy for the P V product
// S [rows = B*H*T, T]; row r belongs to utterance b = r / (H*T); keys j >= lens[b] are padding (-inf): P = softmax_j(S) over the valid keys.
// Pd (optional, f32 or bf16): dropout(P) with the mask of element r*T + j.
__global__ void __launch_bounds__(256) softmax_mask_fwd_kernel(float* __restrict__ S, void* __restrict__ Pd, int pd_dtype, const int32_t* __restrict__ lens,
                                                               long long rows, int HT, int T, float p, uint64_t seed, const unsigned long long* __restrict__ seed_off) {
    if (seed_off) seed += *seed_off;
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int len = lens ? min(T, max(0, lens[r / HT])) : T;
    float* s = S + r * T;
    float mx = -INFINITY;
    for (int j = lane; j < len; j += 32) mx = fmaxf(mx, s[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < len; j += 32) sum += __expf(s[j] - mx);
    sum = warp_sum(sum);
    const float inv = len > 0 ? 1.0f / sum : 0.f, inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.f;
    const uint32_t th = dropout_threshold(p);
    for (int j = lane; j < T; j += 32) {
        const float v = j < len ? __expf(s[j] - mx) * inv : 0.f;
        s[j] = v;
        if (Pd) {
            float d = v;
            if (p > 0.f) {
                const size_t e = (size_t)r * T + j;
                const uint4 rb = dropout_bits(e >> 2, seed);
                const uint32_t w = (e & 3) == 0 ? rb.x : ((e & 3) == 1 ? rb.y : ((e & 3) == 2 ? rb.z : rb.w));
                d = w >= th ? v * inv_keep : 0.f;
            }
            if (pd_dtype == NSD_F32) reinterpret_cast<float*>(Pd)[(size_t)r * T + j] = d;
            else reinterpret_cast<__nv_bfloat16*>(Pd)[(size_t)r * T + j] = __float2bfloat16_rn(d);
        }
    }
}
// dS = P * (dP - sum_j dP_j P_j), dP = dropout-backward of dPd (in place over dPd)
__global__ void __launch_bounds__(256) softmax_mask_bwd_kernel(const float* __restrict__ P, float* __restrict__ dPd, long long rows, int T, float p,
                                                               uint64_t seed, const unsigned long long* __restrict__ seed_off) {
    if (seed_off) seed += *seed_off;
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.f;
    const uint32_t th = dropout_threshold(p);
    float dot = 0.f;
    for (int j = lane; j < T; j += 32) {
        float d = dPd[r * T + j];
        if (p > 0.f) {
            const size_t e = (size_t)r * T + j;
            const uint4 rb = dropout_bits(e >> 2, seed);
            const uint32_t w = (e & 3) == 0 ? rb.x : ((e & 3) == 1 ? rb.y : ((e & 3) == 2 ? rb.z : rb.w));
            d = w >= th ? d * inv_keep : 0.f;
            dPd[r * T + j] = d;
        }
        dot = fmaf(d, P[r * T + j], dot);
    }
    dot = warp_sum(dot);
    for (int j = lane; j < T; j += 32) dPd[r * T + j] = P[r * T + j] * (dPd[r * T + j] - dot);
}

// ---- sum of squares of a list of tensors (clip_grad_norm_, trainer:255-257): per-CTA partials in a fixed order, then one CTA
constexpr int SQ_MAX_TENSORS = 160, SQ_CHUNK = 32768;
struct SqTable { const float* g[SQ_MAX_TENSORS]; long long n[SQ_MAX_TENSORS]; int chunk_start[SQ_MAX_TENSORS + 1]; int count; };
__global__ void __launch_bounds__(256) sqnorm_partial_kernel(const __grid_constant__ SqTable tab, float* __restrict__ part, int part0) {
    __shared__ float sm[8];
    int ti = 0;
    while (ti + 1 < tab.count && (int)blockIdx.x >= tab.chunk_start[ti + 1]) ++ti;
    const long long base = (long long)(blockIdx.x - tab.chunk_start[ti]) * SQ_CHUNK, end = min(tab.n[ti], base + SQ_CHUNK);
    const float* G = tab.g[ti];
    float a = 0.f;
    for (long long i = base + threadIdx.x; i < end; i += 256) a = fmaf(G[i], G[i], a);
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sm[w];
        part[part0 + blockIdx.x] = s;
    }
}
__global__ void __launch_bounds__(1024) sqnorm_final_kernel(const float* __restrict__ part, int n, float* __restrict__ out) {
    __shared__ float sm[32];
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += 1024) a += part[i];
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        a = warp_sum(sm[threadIdx.x]);
        if (threadIdx.x == 0) *out = a;
    }
}

}  // namespace nsd

extern "C" {

using namespace nsd;

int nsd_bgemm(const void* A, int a_dtype, int64_t a_rs, int64_t a_cs, int64_t a_b0, int64_t a_b1, const void* B, int b_dtype, int64_t b_rs, int64_t b_cs,
              int64_t b_b0, int64_t b_b1, const int64_t* b_index, void* C, int c_dtype, int64_t c_rs, int64_t c_b0, int64_t c_b1, const float* bias,
              int64_t bias_b0, int M, int N, int K, int nb0, int nb1, float alpha, int tc_mode, void* stream) {
    NSD_CHECK_ARG(A && B && C && M >= 0 && N >= 0 && K > 0 && nb0 >= 0 && nb1 >= 1, "bgemm: bad argument M=%d N=%d K=%d batches %d x %d", M, N, K, nb0, nb1);
    NSD_CHECK_ARG((a_dtype == NSD_F32 || a_dtype == NSD_BF16) && (b_dtype == NSD_F32 || b_dtype == NSD_BF16) && (c_dtype == NSD_F32 || c_dtype == NSD_BF16),
                  "bgemm: bad dtype");
    if (M == 0 || N == 0 || nb0 == 0) return NSD_OK;
    NSD_CHECK_ARG((long long)nb0 * nb1 <= 65535, "bgemm: %lld batch entries exceed the grid limit", (long long)nb0 * nb1);
    BgemmParams p;
    p.A = A; p.B = B; p.C = C; p.bias = bias; p.b_index = b_index;
    p.a_rs = a_rs; p.a_cs = a_cs; p.a_b0 = a_b0; p.a_b1 = a_b1; p.b_rs = b_rs; p.b_cs = b_cs; p.b_b0 = b_b0; p.b_b1 = b_b1;
    p.c_rs = c_rs; p.c_b0 = c_b0; p.c_b1 = c_b1; p.bias_b0 = bias_b0;
    p.a_dtype = a_dtype; p.b_dtype = b_dtype; p.c_dtype = c_dtype; p.M = M; p.N = N; p.K = K; p.nb1 = nb1; p.alpha = alpha;
    const dim3 grid(cdiv(N, BG_T), cdiv(M, BG_T), nb0 * nb1);
    // bf16 anywhere among the operands -> tensor-core path (operands rounded to bf16 in shared memory); all-fp32 -> FFMA parity path
    const bool tc = tc_mode < 0 ? (a_dtype == NSD_BF16 || b_dtype == NSD_BF16) : tc_mode != 0;
    if (tc) bgemm_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(p);
    else bgemm_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(p);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}

int nsd_softmax_mask_fwd(float* S, void* Pd, int pd_dtype, const int32_t* lens, int B, int H, int T, float p_drop, uint64_t seed, void* stream) {
    NSD_CHECK_ARG(S && B >= 0 && H >= 1 && T >= 1 && p_drop >= 0.f && p_drop < 1.f && (!Pd || pd_dtype == NSD_F32 || pd_dtype == NSD_BF16), "softmax_mask_fwd: bad argument");
    const long long rows = (long long)B * H * T;
    if (rows == 0) return NSD_OK;
    softmax_mask_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(S, Pd, pd_dtype, lens, rows, H * T, T, p_drop, seed, seed_offset_ptr());
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}
int nsd_softmax_mask_bwd(const float* P, float* dPd, int B, int H, int T, float p_drop, uint64_t seed, void* stream) {
    NSD_CHECK_ARG(P && dPd && B >= 0 && H >= 1 && T >= 1 && p_drop >= 0.f && p_drop < 1.f, "softmax_mask_bwd: bad argument");
    const long long rows = (long long)B * H * T;
    if (rows == 0) return NSD_OK;
    softmax_mask_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(P, dPd, rows, T, p_drop, seed, seed_offset_ptr());
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}

size_t nsd_sqnorm_workspace(int n_tensors, const int64_t* numel) {
    size_t chunks = 0;
    for (int i = 0; i < n_tensors; ++i) chunks += cdivz((size_t)std::max<int64_t>(numel[i], 0), SQ_CHUNK);
    return sizeof(float) * (chunks + 1);
}
int nsd_sqnorm_multi(int n_tensors, const void* const* grads, const int64_t* numel, float* out, void* workspace, size_t workspace_bytes, void* stream) {
    NSD_CHECK_ARG(n_tensors >= 0 && out && (n_tensors == 0 || (grads && numel)), "sqnorm_multi: bad argument");
    if (!workspace || workspace_bytes < nsd_sqnorm_workspace(n_tensors, numel)) { set_error("sqnorm_multi: workspace too small"); return NSD_ERR_WORKSPACE; }
    float* part = (float*)workspace;
    int total = 0;
    for (int t0 = 0; t0 < n_tensors; t0 += SQ_MAX_TENSORS) {
        SqTable tab;
        tab.count = std::min(SQ_MAX_TENSORS, n_tensors - t0);
        int chunks = 0;
        for (int i = 0; i < tab.count; ++i) {
            NSD_CHECK_ARG(numel[t0 + i] >= 0 && (numel[t0 + i] == 0 || grads[t0 + i]), "sqnorm_multi: null tensor %d", t0 + i);
            tab.g[i] = (const float*)grads[t0 + i]; tab.n[i] = numel[t0 + i];
            tab.chunk_start[i] = chunks;
            chunks += (int)cdivz((size_t)numel[t0 + i], SQ_CHUNK);
        }
        tab.chunk_start[tab.count] = chunks;
        if (chunks == 0) continue;
        sqnorm_partial_kernel<<<chunks, 256, 0, (cudaStream_t)stream>>>(tab, part, total);
        NSD_LAUNCH_CHECK();
        total += chunks;
    }
    sqnorm_final_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(part, total, out);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}

}  // extern "C"
