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

// loads a [64 rows x 32 k] tile of an operand addressed as base[row*rs + k*cs] into smem[row][k] (zero outside).  Tensor-core path:
// 16-byte global loads along whichever direction is contiguous in the source (8 bf16 / 4 f32 per load; chunks that straddle the
// matrix edge or are misaligned fall back to element loads), so Q, K, V, P and dS tiles cost 2-4 load instructions per thread instead
// of 16.  FFMA (fp32 parity) path: element loads, threads walking the contiguous direction.
struct BgTileSrc {                       // how one operand's [64 x 32] tiles are fetched (fixed for the kernel)
    const void* base; long long rs, cs; int dtype, nrows, K, V, nc, no, nc_sh, no_sh, v_sh, per_thread; bool along_k, vec_ok;
};
__device__ __forceinline__ BgTileSrc bg_src(const void* base, int dtype, long long rs, long long cs, int row0, int nrows, int K) {
    BgTileSrc t;
    t.base = base; t.rs = rs; t.cs = cs; t.dtype = dtype; t.nrows = nrows; t.K = K;
    t.along_k = (cs == 1) || (rs != 1);
    t.V = dtype == NSD_F32 ? 4 : 8;
    const size_t es = dtype == NSD_F32 ? 4 : 2;
    const long long ostr = t.along_k ? rs : cs;
    // every tile starts at (row0, multiple of 32): alignment of the first one decides for all
    t.vec_ok = (t.along_k ? cs == 1 : rs == 1) && (ostr % t.V) == 0 && ((reinterpret_cast<uintptr_t>(base) + (size_t)((long long)row0 * rs) * es) & 15) == 0;
    t.nc = t.along_k ? BG_K / t.V : BG_T / t.V;
    t.no = t.along_k ? BG_T : BG_K;
    t.nc_sh = 31 - __clz(t.nc); t.no_sh = 31 - __clz(t.no); t.v_sh = 31 - __clz(t.V);      // all powers of two: shifts, no divisions
    t.per_thread = t.nc * t.no / 128;            // 2 (bf16) or 4 (f32)
    return t;
}
__device__ __forceinline__ void bg_chunk_coords(const BgTileSrc& t, int c, int& r, int& k) {
    // consecutive threads take consecutive positions of the OTHER direction when the chunks run along the rows (their shared-memory stores
    // then fall into consecutive k: conflict-free), and consecutive chunks of a row when they run along k
    const int ci = t.along_k ? (c & (t.nc - 1)) : (c >> t.no_sh), oi = t.along_k ? (c >> t.nc_sh) : (c & (t.no - 1));
    r = t.along_k ? oi : (ci << t.v_sh);
    k = t.along_k ? (ci << t.v_sh) : oi;
}
// Per-thread chunk state, computed once per CTA: where each of the thread's (up to 4) 16-byte chunks of a k-tile comes from and goes to.
// The k-loop then costs one pointer add + one predicated 16-byte load per chunk (chunks at the matrix edge take the element-wise path).
struct BgThread {
    const char* gp[4];        // address of the chunk in the first k-tile
    int r[4], k[4];           // tile coordinates
    bool fixed_ok[4];         // fast path possible as far as the row direction is concerned
};
__device__ __forceinline__ void bg_thread_init(const BgTileSrc& t, int row0, int tid, BgThread& th) {
    const size_t es = t.dtype == NSD_F32 ? 4 : 2;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        th.gp[q] = nullptr; th.r[q] = th.k[q] = 0; th.fixed_ok[q] = false;
        if (q < t.per_thread) {
            bg_chunk_coords(t, tid + 128 * q, th.r[q], th.k[q]);
            th.gp[q] = reinterpret_cast<const char*>(t.base) + (size_t)((long long)(row0 + th.r[q]) * t.rs + (long long)th.k[q] * t.cs) * es;
            th.fixed_ok[q] = t.vec_ok && (t.along_k ? row0 + th.r[q] < t.nrows : row0 + th.r[q] + t.V <= t.nrows);
        }
    }
}
// element-wise fetch of one chunk that touches the matrix edge (or of a misaligned operand); zeros outside
__device__ __noinline__ uint4 bg_fetch_slow(const BgTileSrc t, int row0, int k0, int r, int k) {
    const int lim = t.along_k ? t.K - (k0 + k) : t.nrows - (row0 + r);
    const bool other_ok = t.along_k ? row0 + r < t.nrows : k0 + k < t.K;
    const int n = other_ok ? max(0, min(t.V, lim)) : 0;
    const long long e0 = (long long)(row0 + r) * t.rs + (long long)(k0 + k) * t.cs, st = t.along_k ? t.cs : t.rs;
    if (t.dtype == NSD_F32) {
        float f[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < n) f[i] = reinterpret_cast<const float*>(t.base)[e0 + (long long)i * st];
        return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    }
    unsigned short h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (i < n) h[i] = reinterpret_cast<const unsigned short*>(t.base)[e0 + (long long)i * st];
    return make_uint4(h[0] | ((uint32_t)h[1] << 16), h[2] | ((uint32_t)h[3] << 16), h[4] | ((uint32_t)h[5] << 16), h[6] | ((uint32_t)h[7] << 16));
}
// request this thread's chunks of the tile at (row0, k0) into registers (raw 16 bytes per chunk)
__device__ __forceinline__ void bg_fetch(const BgTileSrc& t, const BgThread& th, int row0, int k0, size_t kbytes, uint4 (&raw)[4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (q < t.per_thread) {
            const bool k_ok = t.along_k ? k0 + th.k[q] + t.V <= t.K : k0 + th.k[q] < t.K;
            if (th.fixed_ok[q] && k_ok) raw[q] = *reinterpret_cast<const uint4*>(th.gp[q] + kbytes);
            else raw[q] = bg_fetch_slow(t, row0, k0, th.r[q], th.k[q]);
        }
    }
}
// registers -> shared memory [row][k] as bf16 (elements beyond the matrix edge were fetched as zeros)
__device__ __forceinline__ void bg_commit(const BgTileSrc& t, const BgThread& th, __nv_bfloat16 (*sm)[BG_K + BG_PAD], const uint4 (&raw)[4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (q < t.per_thread) {
            const int r = th.r[q], k = th.k[q];
            if (t.dtype == NSD_F32) {
                const float f[4] = {__uint_as_float(raw[q].x), __uint_as_float(raw[q].y), __uint_as_float(raw[q].z), __uint_as_float(raw[q].w)};
                if (t.along_k) {
                    const __nv_bfloat162 lo = __floats2bfloat162_rn(f[0], f[1]), hi = __floats2bfloat162_rn(f[2], f[3]);
                    *reinterpret_cast<uint2*>(&sm[r][k]) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) sm[r + i][k] = __float2bfloat16_rn(f[i]);
                }
            } else if (t.along_k) {
                *reinterpret_cast<uint4*>(&sm[r][k]) = raw[q];
            } else {
                const uint32_t w[4] = {raw[q].x, raw[q].y, raw[q].z, raw[q].w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const unsigned short h = (unsigned short)(w[i >> 1] >> (16 * (i & 1)));
                    sm[r + i][k] = *reinterpret_cast<const __nv_bfloat16*>(&h);
                }
            }
        }
    }
}

template <typename ST>
__device__ __forceinline__ void bg_load_tile(ST (*sm)[BG_K + BG_PAD], const void* base, int dtype, long long rs, long long cs, int row0, int nrows, int k0,
                                             int K, int tid) {
    {
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
}

template <bool TC>
__global__ void __launch_bounds__(128) bgemm_kernel(const BgemmParams p) {
    pdl_enter();
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
    // tensor-core path: the next k-tile of both operands is requested into registers before the MMAs of the current one are issued
    BgTileSrc ta, tb;
    BgThread tha, thb;
    uint4 ra[4], rb[4];
    size_t ka_bytes = 0, kb_bytes = 0;                        // byte step of one k-tile along each operand
    if constexpr (TC) {
        ta = bg_src(A, p.a_dtype, p.a_rs, p.a_cs, m0, p.M, p.K);
        tb = bg_src(Bm, p.b_dtype, p.b_cs, p.b_rs, n0, p.N, p.K);                         // row of Bs = n: (rs, cs) seen from n are (b_cs, b_rs)
        bg_thread_init(ta, m0, tid, tha);
        bg_thread_init(tb, n0, tid, thb);
        ka_bytes = (size_t)BG_K * p.a_cs * ea; kb_bytes = (size_t)BG_K * p.b_rs * eb;
#pragma unroll
        for (int q = 0; q < 4; ++q) ra[q] = rb[q] = make_uint4(0u, 0u, 0u, 0u);
        bg_fetch(ta, tha, m0, 0, 0, ra);
        bg_fetch(tb, thb, n0, 0, 0, rb);
    }
    for (int k0 = 0; k0 < p.K; k0 += BG_K) {
        if constexpr (TC) {
            bg_commit(ta, tha, As, ra);
            bg_commit(tb, thb, Bs, rb);
        } else {
            bg_load_tile<ST>(As, A, p.a_dtype, p.a_rs, p.a_cs, m0, p.M, k0, p.K, tid);
            bg_load_tile<ST>(Bs, Bm, p.b_dtype, p.b_cs, p.b_rs, n0, p.N, k0, p.K, tid);
        }
        __syncthreads();
        if constexpr (TC) {
            if (k0 + BG_K < p.K) {
                const size_t it = (size_t)(k0 / BG_K + 1);
                bg_fetch(ta, tha, m0, k0 + BG_K, it * ka_bytes, ra);
                bg_fetch(tb, thb, n0, k0 + BG_K, it * kb_bytes, rb);
            }
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
            for (int h = 0; h < 2; ++h) {
                const int m = m0 + wm + 16 * i + g + 8 * h, n = n0 + wn + 8 * j + 2 * c;
                if (m < p.M && n < p.N) {
                    const float v0 = fmaf(p.alpha, acc[i][j][2 * h], bias ? bias[n] : 0.f);
                    const float v1 = n + 1 < p.N ? fmaf(p.alpha, acc[i][j][2 * h + 1], bias ? bias[n + 1] : 0.f) : 0.f;
                    const long long o = cbase + (long long)m * p.c_rs + n;
                    const bool pair = n + 1 < p.N && (o & 1) == 0;          // two adjacent outputs in one 8- / 4-byte store when aligned
                    if (p.c_dtype == NSD_F32) {
                        float* cp = reinterpret_cast<float*>(C) + o;
                        if (pair) *reinterpret_cast<float2*>(cp) = make_float2(v0, v1);
                        else { cp[0] = v0; if (n + 1 < p.N) cp[1] = v1; }
                    } else {
                        __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(C) + o;
                        if (pair) *reinterpret_cast<__nv_bfloat162*>(cp) = __floats2bfloat162_rn(v0, v1);
                        else { cp[0] = __float2bfloat16_rn(v0); if (n + 1 < p.N) cp[1] = __float2bfloat16_rn(v1); }
                    }
                }
            }
}

// ---- masked row softmax of the attention scores (in place), optional dropped copy for the P V product
// S [rows = B*H*T, T]; row r belongs to utterance b = r / (H*T); keys j >= lens[b] are padding (-inf): P = softmax_j(S) over the valid keys.
// Pd (optional, f32 or bf16): dropout(P) with the mask of element r*T + j.
__global__ void __launch_bounds__(256) softmax_mask_fwd_kernel(float* __restrict__ S, void* __restrict__ Pd, int pd_dtype, const int32_t* __restrict__ lens,
                                                               long long rows, int HT, int T, int ld, float p, uint64_t seed,
                                                               const unsigned long long* __restrict__ seed_off) {
    pdl_enter();
    // a lane owns groups of 4 consecutive scores (ld % 4 == 0): 16-byte accesses, and ONE Philox block per group for the dropout mask
    // (element r*ld + j: the padded index, so that a group never straddles two blocks)
    if (seed_off) seed += *seed_off;
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int len = lens ? min(T, max(0, lens[r / HT])) : T;
    float4* s4 = reinterpret_cast<float4*>(S + r * ld);
    const int n4 = ld >> 2;
    float mx = -INFINITY;
    for (int q = lane; q < n4; q += 32) {
        const float4 v = s4[q];
        const int j = 4 * q;
        if (j < len) mx = fmaxf(mx, v.x);
        if (j + 1 < len) mx = fmaxf(mx, v.y);
        if (j + 2 < len) mx = fmaxf(mx, v.z);
        if (j + 3 < len) mx = fmaxf(mx, v.w);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int q = lane; q < n4; q += 32) {
        const float4 v = s4[q];
        const int j = 4 * q;
        sum += (j < len ? __expf(v.x - mx) : 0.f) + (j + 1 < len ? __expf(v.y - mx) : 0.f) + (j + 2 < len ? __expf(v.z - mx) : 0.f) +
               (j + 3 < len ? __expf(v.w - mx) : 0.f);
    }
    sum = warp_sum(sum);
    const float inv = len > 0 ? 1.0f / sum : 0.f, inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.f;
    const uint32_t th = dropout_threshold(p);
    for (int q = lane; q < n4; q += 32) {                 // the padding columns [T, ld) are written as zeros
        const float4 v = s4[q];
        const int j = 4 * q;
        float4 o;
        o.x = j < len ? __expf(v.x - mx) * inv : 0.f; o.y = j + 1 < len ? __expf(v.y - mx) * inv : 0.f;
        o.z = j + 2 < len ? __expf(v.z - mx) * inv : 0.f; o.w = j + 3 < len ? __expf(v.w - mx) * inv : 0.f;
        s4[q] = o;
        if (Pd) {
            float4 d = o;
            if (p > 0.f) {
                const uint4 rb = dropout_bits(((size_t)r * ld + j) >> 2, seed);
                d.x = rb.x >= th ? d.x * inv_keep : 0.f; d.y = rb.y >= th ? d.y * inv_keep : 0.f;
                d.z = rb.z >= th ? d.z * inv_keep : 0.f; d.w = rb.w >= th ? d.w * inv_keep : 0.f;
            }
            if (pd_dtype == NSD_F32) reinterpret_cast<float4*>(reinterpret_cast<float*>(Pd) + (size_t)r * ld)[q] = d;
            else {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(d.x, d.y), hi = __floats2bfloat162_rn(d.z, d.w);
                reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(Pd) + (size_t)r * ld)[q] =
                    make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
            }
        }
    }
}
// dS = P * (dP - sum_j dP_j P_j), dP = dropout-backward of dPd (in place over dPd)
__global__ void __launch_bounds__(256) softmax_mask_bwd_kernel(const float* __restrict__ P, float* __restrict__ dPd, long long rows, int T, int ld, float p,
                                                               uint64_t seed, const unsigned long long* __restrict__ seed_off) {
    pdl_enter();
    if (seed_off) seed += *seed_off;
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float inv_keep = p > 0.f ? 1.0f / (1.0f - p) : 1.f;
    const uint32_t th = dropout_threshold(p);
    const float4* p4 = reinterpret_cast<const float4*>(P + r * ld);
    float4* d4 = reinterpret_cast<float4*>(dPd + r * ld);
    const int n4 = ld >> 2;
    float dot = 0.f;
    for (int q = lane; q < n4; q += 32) {
        float4 d = d4[q];
        const float4 pv = p4[q];                           // zero in the padding columns and at masked keys
        const int j = 4 * q;                               // the padding columns of dPd were never written: keep them out of the sums
        if (j >= T) d.x = 0.f;
        if (j + 1 >= T) d.y = 0.f;
        if (j + 2 >= T) d.z = 0.f;
        if (j + 3 >= T) d.w = 0.f;
        if (p > 0.f || j + 3 >= T) {
            if (p > 0.f) {
                const uint4 rb = dropout_bits(((size_t)r * ld + j) >> 2, seed);
                d.x = rb.x >= th ? d.x * inv_keep : 0.f; d.y = rb.y >= th ? d.y * inv_keep : 0.f;
                d.z = rb.z >= th ? d.z * inv_keep : 0.f; d.w = rb.w >= th ? d.w * inv_keep : 0.f;
            }
            d4[q] = d;
        }
        dot += (d.x * pv.x + d.y * pv.y) + (d.z * pv.z + d.w * pv.w);
    }
    dot = warp_sum(dot);
    for (int q = lane; q < n4; q += 32) {
        const float4 d = d4[q], pv = p4[q];
        d4[q] = make_float4(pv.x * (d.x - dot), pv.y * (d.y - dot), pv.z * (d.z - dot), pv.w * (d.w - dot));
    }
}

// ---- sum of squares of a list of tensors (clip_grad_norm_, trainer:255-257): per-CTA partials in a fixed order, then one CTA
constexpr int SQ_MAX_TENSORS = 160, SQ_CHUNK = 32768;
struct SqTable { const float* g[SQ_MAX_TENSORS]; long long n[SQ_MAX_TENSORS]; int chunk_start[SQ_MAX_TENSORS + 1]; int count; };
__global__ void __launch_bounds__(256) sqnorm_partial_kernel(const __grid_constant__ SqTable tab, float* __restrict__ part, int part0) {
    pdl_enter();
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
    pdl_enter();
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
    if (tc) nsd::launch_k(bgemm_kernel<true>, grid, 128, 0, (cudaStream_t)stream, p);
    else nsd::launch_k(bgemm_kernel<false>, grid, 128, 0, (cudaStream_t)stream, p);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}

int nsd_softmax_mask_fwd(float* S, void* Pd, int pd_dtype, const int32_t* lens, int B, int H, int T, int ld, float p_drop, uint64_t seed, void* stream) {
    NSD_CHECK_ARG(S && B >= 0 && H >= 1 && T >= 1 && ld >= T && ld % 4 == 0 && p_drop >= 0.f && p_drop < 1.f && (!Pd || pd_dtype == NSD_F32 || pd_dtype == NSD_BF16), "softmax_mask_fwd: bad argument");
    const long long rows = (long long)B * H * T;
    if (rows == 0) return NSD_OK;
    nsd::launch_k(softmax_mask_fwd_kernel, (unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream, S, Pd, pd_dtype, lens, rows, H * T, T, ld, p_drop, seed, seed_offset_ptr());
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}
int nsd_softmax_mask_bwd(const float* P, float* dPd, int B, int H, int T, int ld, float p_drop, uint64_t seed, void* stream) {
    NSD_CHECK_ARG(P && dPd && B >= 0 && H >= 1 && T >= 1 && ld >= T && ld % 4 == 0 && p_drop >= 0.f && p_drop < 1.f, "softmax_mask_bwd: bad argument");
    const long long rows = (long long)B * H * T;
    if (rows == 0) return NSD_OK;
    nsd::launch_k(softmax_mask_bwd_kernel, (unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream, P, dPd, rows, T, ld, p_drop, seed, seed_offset_ptr());
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
        nsd::launch_k(sqnorm_partial_kernel, chunks, 256, 0, (cudaStream_t)stream, tab, part, total);
        NSD_LAUNCH_CHECK();
        total += chunks;
    }
    nsd::launch_k(sqnorm_final_kernel, 1, 1024, 0, (cudaStream_t)stream, part, total, out);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}

}  // extern "C"
