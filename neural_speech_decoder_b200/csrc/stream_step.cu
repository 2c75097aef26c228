// Streaming inference step for small batches (BASELINE configs[3]): ONE launch advances the whole
// unidirectional GRU stack by one output frame -- every layer's input projection, recurrent projection, gate math and
// state update (reference nn.GRU, model.py:50-57, 104-119, with the state CARRIED between calls, which the reference's
// forward cannot do), the phoneme-logit projection (model.py:122) and the greedy argmax (trainer:314).
//
// At batch 1..8 a step is pure weight bandwidth: 107 MB of bf16 weights (W_ih 50 MB for layer 0, 6.3 MB for each of the other
// nine matrices) against 2 x 107 M multiply-adds per batch row.  The time-batched tcgen05 path pays, per 80 ms push, five
// GEMM launches, five cooperative cluster launches that each restage 6.3 MB of W_hh into tensor memory, and the host
// enqueue of ~15 calls.  Here a persistent cooperative grid streams each weight row exactly once with 16-byte loads (the
// whole set fits the 126 MB L2, so consecutive pushes find it there: the weights stay on chip across calls), one warp per
// hidden unit computes that unit's six dot products for all batch rows, and the layers are separated by grid barriers.
// bf16 weights and bf16-rounded inputs / recurrent state, fp32 accumulation and fp32 carried state: the same arithmetic
// as the tcgen05 path up to summation order.
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace nsd {

constexpr int SS_MAX_LAYERS = 8;
constexpr int SS_THREADS = 256;

struct StreamStepParams {
    const __nv_bfloat16* x0; int ldx;                     // [B, F0] layer-0 input rows (patches of the new frame)
    const __nv_bfloat16* w_ih[SS_MAX_LAYERS]; const __nv_bfloat16* w_hh[SS_MAX_LAYERS];   // [3H, in_l], [3H, H], gate rows r | z | n
    const float* b_ih[SS_MAX_LAYERS]; const float* b_hh[SS_MAX_LAYERS];                   // [3H]
    float* h;                                             // [L][B][H] carried state, updated in place
    float* h_new; __nv_bfloat16* h_new_bf;                // workspace: [L][B][H] each
    const __nv_bfloat16* fc_w; const float* fc_b;         // [C, H], [C]
    float* logits; int* ids;                              // [B, C], [B]
    int B, F0, H, L, C;
};

__device__ __forceinline__ float ss_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ss_sigmoid(float x) { return fmaf(0.5f, ss_tanh(0.5f * x), 0.5f); }

__device__ __forceinline__ void bf16x8_to_f32(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 v = __bfloat1622float2(h[i]);
        f[2 * i] = v.x; f[2 * i + 1] = v.y;
    }
}
__device__ __forceinline__ float dot8(const uint4& w, const float (&x)[8], float acc) {
    float wf[8];
    bf16x8_to_f32(w, wf);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(wf[i], x[i], acc);
    return acc;
}

template <int BM>
__global__ void __launch_bounds__(SS_THREADS) stream_step_kernel(const StreamStepParams p) {
    cg::grid_group grid = cg::this_grid();
    const int H = p.H, B = p.B;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * (SS_THREADS / 32) + warp, nw = gridDim.x * (SS_THREADS / 32);
    for (int l = 0; l < p.L; ++l) {
        const int in_l = l == 0 ? p.F0 : H;
        const __nv_bfloat16* xin = l == 0 ? p.x0 : p.h_new_bf + (size_t)(l - 1) * B * H;
        const int ldx = l == 0 ? p.ldx : H;
        const float* hp = p.h + (size_t)l * B * H;
        const __nv_bfloat16* Wi = p.w_ih[l];
        const __nv_bfloat16* Wh = p.w_hh[l];
        for (int u = gw; u < H; u += nw) {
            float ai[3][BM], ah[3][BM];
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int b = 0; b < BM; ++b) ai[g][b] = ah[g][b] = 0.f;
            // input projection: rows u, H+u, 2H+u of W_ih against the layer input
            for (int k0 = lane * 8; k0 < in_l; k0 += 256) {
                const uint4 wr = __ldg(reinterpret_cast<const uint4*>(Wi + (size_t)u * in_l + k0));
                const uint4 wz = __ldg(reinterpret_cast<const uint4*>(Wi + (size_t)(H + u) * in_l + k0));
                const uint4 wn = __ldg(reinterpret_cast<const uint4*>(Wi + (size_t)(2 * H + u) * in_l + k0));
#pragma unroll
                for (int b = 0; b < BM; ++b) {
                    if (b < B) {
                        float xv[8];
                        bf16x8_to_f32(*reinterpret_cast<const uint4*>(xin + (size_t)b * ldx + k0), xv);   // written by this launch for l > 0: plain load
                        ai[0][b] = dot8(wr, xv, ai[0][b]); ai[1][b] = dot8(wz, xv, ai[1][b]); ai[2][b] = dot8(wn, xv, ai[2][b]);
                    }
                }
            }
            // recurrent projection: the previous state enters as bf16, like the B operand of the tcgen05 recurrence
            for (int k0 = lane * 8; k0 < H; k0 += 256) {
                const uint4 wr = __ldg(reinterpret_cast<const uint4*>(Wh + (size_t)u * H + k0));
                const uint4 wz = __ldg(reinterpret_cast<const uint4*>(Wh + (size_t)(H + u) * H + k0));
                const uint4 wn = __ldg(reinterpret_cast<const uint4*>(Wh + (size_t)(2 * H + u) * H + k0));
#pragma unroll
                for (int b = 0; b < BM; ++b) {
                    if (b < B) {
                        const float4 h0 = *reinterpret_cast<const float4*>(hp + (size_t)b * H + k0), h1 = *reinterpret_cast<const float4*>(hp + (size_t)b * H + k0 + 4);
                        float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i) hv[i] = __bfloat162float(__float2bfloat16_rn(hv[i]));
                        ah[0][b] = dot8(wr, hv, ah[0][b]); ah[1][b] = dot8(wz, hv, ah[1][b]); ah[2][b] = dot8(wn, hv, ah[2][b]);
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int b = 0; b < BM; ++b) { ai[g][b] = warp_sum(ai[g][b]); ah[g][b] = warp_sum(ah[g][b]); }
            if (lane == 0) {
                const float bir = p.b_ih[l][u], biz = p.b_ih[l][H + u], bin = p.b_ih[l][2 * H + u];
                const float bhr = p.b_hh[l][u], bhz = p.b_hh[l][H + u], bhn = p.b_hh[l][2 * H + u];
#pragma unroll
                for (int b = 0; b < BM; ++b) {
                    if (b < B) {
                        const float r = ss_sigmoid((ai[0][b] + bir) + bhr + ah[0][b]);
                        const float z = ss_sigmoid((ai[1][b] + biz) + bhz + ah[1][b]);
                        const float gn = ah[2][b] + bhn;
                        const float n = ss_tanh(fmaf(r, gn, ai[2][b] + bin));
                        const float hprev = hp[(size_t)b * H + u];
                        const float hn = fmaf(z, hprev - n, n);                    // (1-z)*n + z*h_prev
                        p.h_new[((size_t)l * B + b) * H + u] = hn;
                        p.h_new_bf[((size_t)l * B + b) * H + u] = __float2bfloat16_rn(hn);
                    }
                }
            }
        }
        grid.sync();
    }
    // carried state <- new state (every CTA copies a slice), logits + greedy id by block 0
    const size_t nstate = (size_t)p.L * B * H;
    for (size_t i = (size_t)blockIdx.x * SS_THREADS + threadIdx.x; i < nstate; i += (size_t)gridDim.x * SS_THREADS) p.h[i] = p.h_new[i];
    if (blockIdx.x == 0) {
        extern __shared__ float lg[];                      // [B][C]
        const __nv_bfloat16* top = p.h_new_bf + (size_t)(p.L - 1) * B * H;
        for (int o = warp; o < B * p.C; o += SS_THREADS / 32) {
            const int b = o / p.C, c = o - b * p.C;
            float acc = 0.f;
            for (int k0 = lane * 8; k0 < H; k0 += 256) {
                float xv[8];
                bf16x8_to_f32(*reinterpret_cast<const uint4*>(top + (size_t)b * H + k0), xv);
                acc = dot8(__ldg(reinterpret_cast<const uint4*>(p.fc_w + (size_t)c * H + k0)), xv, acc);
            }
            acc = warp_sum(acc);
            if (lane == 0) { acc += p.fc_b[c]; lg[o] = acc; p.logits[o] = acc; }
        }
        __syncthreads();
        if (p.ids != nullptr && threadIdx.x < B) {         // argmax, ties -> lowest index (trainer:314: torch.argmax)
            const float* row = lg + threadIdx.x * p.C;
            int best = 0;
            for (int c = 1; c < p.C; ++c)
                if (row[c] > row[best]) best = c;
            p.ids[threadIdx.x] = best;
        }
    }
}

}  // namespace nsd

extern "C" {

size_t nsd_gru_stream_step_workspace(int B, int H, int L) { return (size_t)L * B * H * (sizeof(float) + sizeof(__nv_bfloat16)) + 256; }

int nsd_gru_stream_step(const void* x0_bf16, int ldx, int B, int F0, int H, int L, int C, const void* const* w_ih_bf16,
                        const void* const* w_hh_bf16, const void* const* b_ih, const void* const* b_hh, float* h,
                        const void* fc_w_bf16, const float* fc_b, float* logits, int* ids, void* workspace, size_t workspace_bytes,
                        void* stream) {
    using namespace nsd;
    NSD_CHECK_ARG(B >= 1 && B <= 8, "gru_stream_step: batch %d not in [1, 8] (larger batches take the time-batched path)", B);
    NSD_CHECK_ARG(L >= 1 && L <= SS_MAX_LAYERS && H > 0 && H % 256 == 0 && F0 > 0 && F0 % 256 == 0 && C > 0 && (ldx % 8) == 0,
                  "gru_stream_step: bad sizes L=%d H=%d F0=%d C=%d ldx=%d (H and F0 must be multiples of 256)", L, H, F0, C, ldx);
    NSD_CHECK_ARG(x0_bf16 && w_ih_bf16 && w_hh_bf16 && b_ih && b_hh && h && fc_w_bf16 && fc_b && logits && workspace, "gru_stream_step: null pointer");
    if (workspace_bytes < nsd_gru_stream_step_workspace(B, H, L)) { set_error("gru_stream_step: workspace too small"); return NSD_ERR_WORKSPACE; }
    StreamStepParams p;
    p.x0 = reinterpret_cast<const __nv_bfloat16*>(x0_bf16); p.ldx = ldx;
    for (int l = 0; l < L; ++l) {
        NSD_CHECK_ARG(w_ih_bf16[l] && w_hh_bf16[l] && b_ih[l] && b_hh[l], "gru_stream_step: null weight pointer for layer %d", l);
        p.w_ih[l] = reinterpret_cast<const __nv_bfloat16*>(w_ih_bf16[l]); p.w_hh[l] = reinterpret_cast<const __nv_bfloat16*>(w_hh_bf16[l]);
        p.b_ih[l] = reinterpret_cast<const float*>(b_ih[l]); p.b_hh[l] = reinterpret_cast<const float*>(b_hh[l]);
    }
    p.h = h;
    p.h_new = reinterpret_cast<float*>(workspace);
    p.h_new_bf = reinterpret_cast<__nv_bfloat16*>(p.h_new + (size_t)L * B * H);
    p.fc_w = reinterpret_cast<const __nv_bfloat16*>(fc_w_bf16); p.fc_b = fc_b; p.logits = logits; p.ids = ids;
    p.B = B; p.F0 = F0; p.H = H; p.L = L; p.C = C;
    const size_t smem = sizeof(float) * (size_t)B * C;
    auto go = [&](auto kern) -> int {
        int per_sm = 0;
        NSD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SS_THREADS, smem));
        if (per_sm < 1) { set_error("gru_stream_step: kernel does not fit an SM"); return NSD_ERR_INVALID; }
        const int grid = sm_count() * std::min(per_sm, 2);                 // every block co-resident (grid barrier between layers)
        void* args[] = {(void*)&p};
        NSD_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(SS_THREADS), args, smem, (cudaStream_t)stream));
        count_launch(1);
        return NSD_OK;
    };
    if (B == 1) return go(stream_step_kernel<1>);
    if (B == 2) return go(stream_step_kernel<2>);
    if (B <= 4) return go(stream_step_kernel<4>);
    return go(stream_step_kernel<8>);
}

}  // extern "C"
