// fp32 CUDA-core GEMM (the fp32 parity path of K2) plus small helpers (column sums, casts, [T,B,C]<->[B,T,C]).
//
// C[M,N] = op(A)[M,K] * op(B)[K,N] (+bias) (+beta*C), row-major, fp32 FFMA with a 128x128x16 CTA tile,
// 8x8 register tile per thread and a double-buffered shared-memory pipeline.  It backs the time-batched
// W_ih projections, their dgrad/wgrad and the output layer in fp32 mode (nn.GRU / nn.Linear in the
// reference, model.py:50-57, 76-81, 119, 122); the tensor-core path lives in gemm_tc.cu.
#include "common.cuh"

namespace nsd {

constexpr int GS_BM = 128, GS_BN = 128, GS_BK = 16, GS_THREADS = 256;

// Loads a [rows=GS_BK (k)] x [cols=128 (m or n)] tile into registers.  `kcontig` selects which logical
// dimension is contiguous in memory.
//   kcontig:   element (mn, k) at base[mn*ld + k]
//   !kcontig:  element (mn, k) at base[k*ld + mn]
template <bool KCONTIG>
struct TileLoader {
    const float* base; int ld, mn0, mn_lim, k_lim; bool vec;
    // each thread owns 8 elements of the 128x16 tile
    __device__ __forceinline__ void load(int k0, int tid, float (&r)[8]) const {
        if constexpr (KCONTIG) {
            // 128 rows(mn) x 16 k: thread -> row = tid/2, k-half = (tid&1)*8
            const int mn = mn0 + (tid >> 1), kk = k0 + ((tid & 1) << 3);
            if (mn < mn_lim && vec && kk + 8 <= k_lim) {
                const float4* p = reinterpret_cast<const float4*>(base + (size_t)mn * ld + kk);
                float4 a = __ldg(p), b = __ldg(p + 1);
                r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = (mn < mn_lim && kk + i < k_lim) ? __ldg(base + (size_t)mn * ld + kk + i) : 0.f;
            }
        } else {
            // 16 k-rows x 128 mn: thread -> k = tid/16, mn group = (tid&15)*8
            const int kk = k0 + (tid >> 4), mn = mn0 + ((tid & 15) << 3);
            if (kk < k_lim && vec && mn + 8 <= mn_lim) {
                const float4* p = reinterpret_cast<const float4*>(base + (size_t)kk * ld + mn);
                float4 a = __ldg(p), b = __ldg(p + 1);
                r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = (kk < k_lim && mn + i < mn_lim) ? __ldg(base + (size_t)kk * ld + mn + i) : 0.f;
            }
        }
    }
    // smem tile layout: s[k][mn] with row length 128 (+4 pad when written transposed)
    __device__ __forceinline__ void store(float* s, int tid, const float (&r)[8]) const {
        if constexpr (KCONTIG) {
            const int mn = tid >> 1, kk = (tid & 1) << 3;
#pragma unroll
            for (int i = 0; i < 8; ++i) s[(kk + i) * (GS_BM + 4) + mn] = r[i];
        } else {
            const int kk = tid >> 4, mn = (tid & 15) << 3;
            float4* d = reinterpret_cast<float4*>(s + kk * (GS_BM + 4) + mn);
            d[0] = make_float4(r[0], r[1], r[2], r[3]);
            d[1] = make_float4(r[4], r[5], r[6], r[7]);
        }
    }
};

template <bool A_KCONTIG, bool B_KCONTIG>
__global__ void __launch_bounds__(GS_THREADS, 2)
sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
             float* __restrict__ C, int ldc, const float* __restrict__ bias, float beta) {
    pdl_enter();
    __shared__ __align__(16) float As[2][GS_BK * (GS_BM + 4)];
    __shared__ __align__(16) float Bs[2][GS_BK * (GS_BN + 4)];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * GS_BM, n0 = blockIdx.x * GS_BN;
    TileLoader<A_KCONTIG> la{A, lda, m0, M, K, false};
    TileLoader<B_KCONTIG> lb{B, ldb, n0, N, K, false};
    la.vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    lb.vec = ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);

    const int ty = tid >> 4, tx = tid & 15;     // 16x16 threads, each 8x8 outputs (split 4+4 for conflict-free LDS.128)
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[8], rb[8];
    la.load(0, tid, ra);
    lb.load(0, tid, rb);
    la.store(As[0], tid, ra);
    lb.store(Bs[0], tid, rb);
    __syncthreads();
    const int nk = (K + GS_BK - 1) / GS_BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            la.load((kt + 1) * GS_BK, tid, ra);
            lb.load((kt + 1) * GS_BK, tid, rb);
        }
        const float* as = As[cur];
        const float* bs = Bs[cur];
#pragma unroll
        for (int k = 0; k < GS_BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(as + k * (GS_BM + 4) + ty * 4);
            float4 a1 = *reinterpret_cast<const float4*>(as + k * (GS_BM + 4) + 64 + ty * 4);
            float4 b0 = *reinterpret_cast<const float4*>(bs + k * (GS_BN + 4) + tx * 4);
            float4 b1 = *reinterpret_cast<const float4*>(bs + k * (GS_BN + 4) + 64 + tx * 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            la.store(As[cur ^ 1], tid, ra);
            lb.store(Bs[cur ^ 1], tid, rb);
        }
        __syncthreads();
    }
    // epilogue: rows ty*4+{0..3} and 64+ty*4+{0..3}; cols tx*4+{0..3} and 64+tx*4+{0..3}
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= N) continue;
            float v = acc[i][j];
            if (bias) v += __ldg(bias + n);
            float* c = C + (size_t)m * ldc + n;
            if (beta != 0.f) v = fmaf(beta, *c, v);
            *c = v;
        }
    }
}

// Column sums (bias gradients), two deterministic stages: partial[c][n] over row chunk c, then a fixed-order sum over c.
// VEC = elements per thread along n (2 for bf16 so that a warp reads full 128-byte rows).
constexpr int CS_CHUNKS = 64;
template <typename T, int VEC>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const T* __restrict__ a, int M, int N, int lda, int rows_per_chunk,
                                                             float* __restrict__ partial) {
    pdl_enter();
    __shared__ float red[4][64 * VEC + 1];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int n = (blockIdx.x * 64 + tx) * VEC;
    const int m0 = blockIdx.y * rows_per_chunk, m1 = min(M, m0 + rows_per_chunk);
    float s[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) s[v] = 0.f;
    if (n < N) {
        for (int m = m0 + ty; m < m1; m += 4) {
            if constexpr (VEC == 2 && sizeof(T) == 2) {
                if (n + 1 < N) {
                    const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(a + (size_t)m * lda + n);
                    s[0] += __bfloat162float(v2.x); s[1] += __bfloat162float(v2.y);
                } else s[0] += to_f32<T>(a[(size_t)m * lda + n]);
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    if (n + v < N) s[v] += to_f32<T>(a[(size_t)m * lda + n + v]);
            }
        }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) red[ty][tx * VEC + v] = s[v];
    __syncthreads();
    if (ty == 0 && n < N) {
#pragma unroll
        for (int v = 0; v < VEC; ++v)
            if (n + v < N) partial[(size_t)blockIdx.y * N + n + v] = (red[0][tx * VEC + v] + red[1][tx * VEC + v]) + (red[2][tx * VEC + v] + red[3][tx * VEC + v]);
    }
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, int N, int chunks, float* __restrict__ out) {
    pdl_enter();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += partial[(size_t)c * N + n];
    out[n] = s;
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, size_t n) {
    pdl_enter();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = from_f32<D>(to_f32<S>(src[i]));
}

// dst[r][c] = bf16(src[r][c]) and/or dstT[c][r] = bf16(src[r][c]); 64x64 tiles through shared memory so that both
// outputs are written in full rows.  Makes the K-major operand copies the tcgen05 GEMM needs.  VEC2: every thread moves
// element PAIRS (8-byte f32 / 4-byte bf16 loads, 4-byte bf16x2 stores in both layouts); needs even leading dimensions,
// 4/8-byte aligned bases; odd tails fall to the scalar instantiation.
template <typename S, bool VEC2>
__global__ void __launch_bounds__(256) cast_transpose_kernel(const S* __restrict__ src, int R, int Cn, int ld_src,
                                                             __nv_bfloat16* __restrict__ dst, int ld_dst,
                                                             __nv_bfloat16* __restrict__ dstT, int ld_dstT) {
    pdl_enter();
    __shared__ float tile[64][65];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    if constexpr (VEC2) {
        for (int i = ty; i < 64; i += 8) {
            const int r = r0 + i, c = c0 + 2 * tx;
            float v0 = 0.f, v1 = 0.f;
            if (r < R && c < Cn) {               // Cn is even here, so c + 1 < Cn too
                if constexpr (sizeof(S) == 4) { const float2 v = *reinterpret_cast<const float2*>(src + (size_t)r * ld_src + c); v0 = v.x; v1 = v.y; }
                else { const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(src + (size_t)r * ld_src + c); v0 = __bfloat162float(v.x); v1 = __bfloat162float(v.y); }
                if (dst) *reinterpret_cast<__nv_bfloat162*>(dst + (size_t)r * ld_dst + c) = __floats2bfloat162_rn(v0, v1);
            }
            tile[i][2 * tx] = v0; tile[i][2 * tx + 1] = v1;
        }
        if (dstT == nullptr) return;
        __syncthreads();
        for (int i = ty; i < 64; i += 8) {
            const int c = c0 + i, r = r0 + 2 * tx;
            if (c < Cn && r < R)                 // R is even here
                *reinterpret_cast<__nv_bfloat162*>(dstT + (size_t)c * ld_dstT + r) = __floats2bfloat162_rn(tile[2 * tx][i], tile[2 * tx + 1][i]);
        }
    } else {
        for (int i = ty; i < 64; i += 8) {
            const int r = r0 + i;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = c0 + tx + 32 * h;
                float v = 0.f;
                if (r < R && c < Cn) {
                    v = to_f32<S>(src[(size_t)r * ld_src + c]);
                    if (dst) dst[(size_t)r * ld_dst + c] = __float2bfloat16_rn(v);
                }
                tile[i][tx + 32 * h] = v;
            }
        }
        if (dstT == nullptr) return;
        __syncthreads();
        for (int i = ty; i < 64; i += 8) {
            const int c = c0 + i;
            if (c >= Cn) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = r0 + tx + 32 * h;
                if (r < R) dstT[(size_t)c * ld_dstT + r] = __float2bfloat16_rn(tile[tx + 32 * h][i]);
            }
        }
    }
}

__global__ void swap01_kernel(const float* __restrict__ in, float* __restrict__ out, int D0, int D1, int C) {
    pdl_enter();
    // out[d1][d0][c] = in[d0][d1][c]
    const size_t total = (size_t)D0 * D1 * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        size_t r = i / C;
        int d0 = (int)(r % D0), d1 = (int)(r / D0);
        out[i] = in[((size_t)d0 * D1 + d1) * C + c];
    }
}

// Batched bf16 transposes of equally shaped matrices in ONE launch (the W_hh^T operands of every layer and direction for the BPTT:
// ten 3072 x 1024 transposes of 6 MB each are ~12 us apiece as separate launches -- too small to fill the GPU -- and ~25 us together).
constexpr int TR_MAX = 64;
struct TransposeTable { const __nv_bfloat16* src[TR_MAX]; __nv_bfloat16* dstT[TR_MAX]; };
__global__ void __launch_bounds__(256) transpose_bf16_multi_kernel(const __grid_constant__ TransposeTable tab, int R, int Cn, int ld_src, int ld_dstT) {
    pdl_enter();
    __shared__ uint32_t tile[64][33];                    // [row][column pair]
    const __nv_bfloat16* __restrict__ src = tab.src[blockIdx.z];
    __nv_bfloat16* __restrict__ dstT = tab.dstT[blockIdx.z];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    for (int i = ty; i < 64; i += 8) {
        const int r = r0 + i, c = c0 + 2 * tx;
        uint32_t v = 0u;
        if (r < R && c < Cn) v = *reinterpret_cast<const uint32_t*>(src + (size_t)r * ld_src + c);      // R, Cn, ld even: c + 1 < Cn too
        tile[i][tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 64; i += 8) {
        const int c = c0 + i, r = r0 + 2 * tx;
        if (c < Cn && r < R) {
            const uint32_t a = tile[2 * tx][i >> 1], b = tile[2 * tx + 1][i >> 1];
            const uint32_t lo = (i & 1) ? (a >> 16) : (a & 0xffffu), hi = (i & 1) ? (b >> 16) : (b & 0xffffu);
            *reinterpret_cast<uint32_t*>(dstT + (size_t)c * ld_dstT + r) = lo | (hi << 16);
        }
    }
}

}  // namespace nsd

extern "C" {

int nsd_gemm_f32(int transa, int transb, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                 float* C, int ldc, const float* bias, float beta, void* stream) {
    using namespace nsd;
    NSD_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "gemm_f32: negative size");
    if (M == 0 || N == 0) return NSD_OK;
    NSD_CHECK_ARG(A && B && C, "gemm_f32: null pointer");
    dim3 grid(cdiv(N, GS_BN), cdiv(M, GS_BM));
    cudaStream_t s = (cudaStream_t)stream;
    // op(A) is k-contiguous when A is stored [M,K]; op(B) is k-contiguous when B is stored [N,K]
    if (!transa && transb) nsd::launch_k(sgemm_kernel<true, true>, grid, GS_THREADS, 0, s, M, N, K, A, lda, B, ldb, C, ldc, bias, beta);
    else if (!transa && !transb) nsd::launch_k(sgemm_kernel<true, false>, grid, GS_THREADS, 0, s, M, N, K, A, lda, B, ldb, C, ldc, bias, beta);
    else if (transa && !transb) nsd::launch_k(sgemm_kernel<false, false>, grid, GS_THREADS, 0, s, M, N, K, A, lda, B, ldb, C, ldc, bias, beta);
    else nsd::launch_k(sgemm_kernel<false, true>, grid, GS_THREADS, 0, s, M, N, K, A, lda, B, ldb, C, ldc, bias, beta);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}

size_t nsd_colsum_workspace(int N) { return sizeof(float) * (size_t)nsd::CS_CHUNKS * (size_t)(N > 0 ? N : 1); }

int nsd_colsum(const void* a, int a_dtype, int M, int N, int lda, float* out, void* workspace, size_t workspace_bytes,
               void* stream) {
    using namespace nsd;
    NSD_CHECK_ARG(M >= 0 && N > 0, "colsum: bad size");
    if (workspace_bytes < nsd_colsum_workspace(N)) { set_error("colsum: workspace too small"); return NSD_ERR_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    float* partial = reinterpret_cast<float*>(workspace);
    const int rows_per_chunk = std::max(4, cdiv(cdiv(std::max(M, 1), CS_CHUNKS), 4) * 4);
    const int chunks = std::max(1, cdiv(M, rows_per_chunk));
    if (a_dtype == NSD_BF16) {
        const bool vec = ((lda & 1) == 0) && (((uintptr_t)a & 3) == 0);
        if (vec) nsd::launch_k(colsum_partial_kernel<__nv_bfloat16, 2>, dim3(cdiv(N, 128), chunks), 256, 0, s, (const __nv_bfloat16*)a, M, N, lda, rows_per_chunk, partial);
        else nsd::launch_k(colsum_partial_kernel<__nv_bfloat16, 1>, dim3(cdiv(N, 64), chunks), 256, 0, s, (const __nv_bfloat16*)a, M, N, lda, rows_per_chunk, partial);
    } else if (a_dtype == NSD_F32) {
        nsd::launch_k(colsum_partial_kernel<float, 1>, dim3(cdiv(N, 64), chunks), 256, 0, s, (const float*)a, M, N, lda, rows_per_chunk, partial);
    } else { set_error("colsum: bad dtype"); return NSD_ERR_INVALID; }
    NSD_LAUNCH_CHECK();
    nsd::launch_k(colsum_final_kernel, cdiv(N, 256), 256, 0, s, partial, N, chunks, out);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}

int nsd_cast(const void* src, int src_dtype, void* dst, int dst_dtype, size_t n, void* stream) {
    using namespace nsd;
    if (n == 0) return NSD_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int blocks = (int)std::min<size_t>(cdivz(n, 256), (size_t)sm_count() * 16);
    if (src_dtype == NSD_F32 && dst_dtype == NSD_BF16) nsd::launch_k(cast_kernel<float, __nv_bfloat16>, blocks, 256, 0, s, (const float*)src, (__nv_bfloat16*)dst, n);
    else if (src_dtype == NSD_BF16 && dst_dtype == NSD_F32) nsd::launch_k(cast_kernel<__nv_bfloat16, float>, blocks, 256, 0, s, (const __nv_bfloat16*)src, (float*)dst, n);
    else if (src_dtype == NSD_F32 && dst_dtype == NSD_F32) nsd::launch_k(cast_kernel<float, float>, blocks, 256, 0, s, (const float*)src, (float*)dst, n);
    else { set_error("cast: unsupported dtype pair %d->%d", src_dtype, dst_dtype); return NSD_ERR_INVALID; }
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}

int nsd_cast_transpose(const void* src, int src_dtype, int R, int Cn, int ld_src, void* dst, int ld_dst, void* dstT,
                       int ld_dstT, void* stream) {
    using namespace nsd;
    NSD_CHECK_ARG(R >= 0 && Cn >= 0 && (dst || dstT), "cast_transpose: bad arguments");
    if (R == 0 || Cn == 0) return NSD_OK;
    dim3 grid(cdiv(Cn, 64), cdiv(R, 64));
    cudaStream_t s = (cudaStream_t)stream;
    const int esz = (src_dtype == NSD_F32) ? 4 : 2;
    const bool vec = (R % 2 == 0) && (Cn % 2 == 0) && (ld_src % 2 == 0) && (((uintptr_t)src & (size_t)(2 * esz - 1)) == 0) &&
                     (!dst || ((ld_dst % 2 == 0) && (((uintptr_t)dst & 3) == 0))) && (!dstT || ((ld_dstT % 2 == 0) && (((uintptr_t)dstT & 3) == 0)));
    __nv_bfloat16* d0 = (__nv_bfloat16*)dst; __nv_bfloat16* d1 = (__nv_bfloat16*)dstT;
    if (src_dtype == NSD_F32) {
        if (vec) nsd::launch_k(cast_transpose_kernel<float, true>, grid, 256, 0, s, (const float*)src, R, Cn, ld_src, d0, ld_dst, d1, ld_dstT);
        else nsd::launch_k(cast_transpose_kernel<float, false>, grid, 256, 0, s, (const float*)src, R, Cn, ld_src, d0, ld_dst, d1, ld_dstT);
    } else if (src_dtype == NSD_BF16) {
        if (vec) nsd::launch_k(cast_transpose_kernel<__nv_bfloat16, true>, grid, 256, 0, s, (const __nv_bfloat16*)src, R, Cn, ld_src, d0, ld_dst, d1, ld_dstT);
        else nsd::launch_k(cast_transpose_kernel<__nv_bfloat16, false>, grid, 256, 0, s, (const __nv_bfloat16*)src, R, Cn, ld_src, d0, ld_dst, d1, ld_dstT);
    } else { set_error("cast_transpose: bad dtype"); return NSD_ERR_INVALID; }
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}

int nsd_transpose_bf16_multi(int n, const void* const* src, void* const* dstT, int R, int Cn, int ld_src, int ld_dstT, void* stream) {
    using namespace nsd;
    NSD_CHECK_ARG(n >= 0 && R >= 0 && Cn >= 0, "transpose_bf16_multi: bad sizes");
    if (n == 0 || R == 0 || Cn == 0) return NSD_OK;
    NSD_CHECK_ARG(R % 2 == 0 && Cn % 2 == 0 && ld_src % 2 == 0 && ld_dstT % 2 == 0 && ld_src >= Cn && ld_dstT >= R,
                  "transpose_bf16_multi: R=%d C=%d ld_src=%d ld_dstT=%d must be even (and ld >= row length)", R, Cn, ld_src, ld_dstT);
    for (int t0 = 0; t0 < n; t0 += TR_MAX) {
        TransposeTable tab;
        const int cnt = std::min(TR_MAX, n - t0);
        for (int i = 0; i < cnt; ++i) {
            NSD_CHECK_ARG(src[t0 + i] && dstT[t0 + i] && (((uintptr_t)src[t0 + i] | (uintptr_t)dstT[t0 + i]) & 3) == 0, "transpose_bf16_multi: matrix %d null or misaligned", t0 + i);
            tab.src[i] = (const __nv_bfloat16*)src[t0 + i]; tab.dstT[i] = (__nv_bfloat16*)dstT[t0 + i];
        }
        nsd::launch_k(transpose_bf16_multi_kernel, dim3(cdiv(Cn, 64), cdiv(R, 64), cnt), 256, 0, (cudaStream_t)stream, tab, R, Cn, ld_src, ld_dstT);
        NSD_LAUNCH_CHECK();
    }
    return NSD_OK;
}

int nsd_swap01_f32(const float* in, float* out, int D0, int D1, int C, void* stream) {
    using namespace nsd;
    const size_t total = (size_t)D0 * D1 * C;
    if (total == 0) return NSD_OK;
    int blocks = (int)std::min<size_t>(cdivz(total, 256), (size_t)sm_count() * 8);
    nsd::launch_k(swap01_kernel, blocks, 256, 0, (cudaStream_t)stream, in, out, D0, D1, C);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}

}  // extern "C"
