// K4: CTC loss forward + gradient in one launch; K5: greedy decode; on-device edit distance.
//
// K4 replaces log_softmax(2).permute(1,0,2) + torch.nn.CTCLoss(blank=0, reduction="mean",
// zero_infinity=True) + its backward (reference neural_decoder_trainer.py:139-141, 210, 213-218, 242, 252).
// One CTA per utterance: warp 0 runs the alpha recursion forward in time while warp 1 runs
// the beta recursion backward in time over the blank-extended label lattice in log space; both warps
// then form the occupancy sums and the gradient.  Every reduction has a fixed order, so the result is
// bit-reproducible run to run.  Lengths and targets stay on the device (the reference syncs them to host).
#include <math.h>

#include "common.cuh"

namespace nsd {

constexpr int CTC_THREADS = 256;     // warps 0/1 run the alpha/beta recursions; all 8 share the per-frame work before and after
constexpr int CTC_WARPS = CTC_THREADS / 32;
#define NSD_NEG_INF (-INFINITY)

__device__ __forceinline__ float lse2f(float a, float b) {
    const float m = fmaxf(a, b);
    if (m == NSD_NEG_INF) return NSD_NEG_INF;
    return m + logf(expf(a - m) + expf(b - m));
}
__device__ __forceinline__ float lse3f(float a, float b, float c) {
    const float m = fmaxf(a, fmaxf(b, c));
    if (m == NSD_NEG_INF) return NSD_NEG_INF;
    return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

struct CtcParams {
    const float* act; int64_t st, sb, sc; int is_logits;
    const int32_t* targets; int tgt_stride; const int32_t* in_lens; const int32_t* tgt_lens;
    int T, B, C, blank, SP, reduction_mean;
    float* nll; float* loss; float* grad;
    float* lp; float* alpha; float* beta; unsigned int* counter;
};

__global__ void __launch_bounds__(CTC_THREADS) ctc_kernel(CtcParams p) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = p.T, C = p.C, SP = p.SP;
    int il = min(max(p.in_lens[b], 0), T);
    int tl = min(max(p.tgt_lens[b], 0), (SP - 1) / 2);
    const int S = 2 * tl + 1;

    int* ext = reinterpret_cast<int*>(smem_raw);                     // [SP]   blank-extended labels
    int* nxt = ext + SP;                                              // [SP]   next lattice slot with the same label (or -1)
    int* isfirst = nxt + SP;                                          // [SP]   1 if no earlier slot carries the same label
    float* rowbuf = reinterpret_cast<float*>(isfirst + SP);           // [2 warps][2][SP+2]  alpha / beta ping-pong
    float* vbuf = rowbuf + 2 * 2 * (SP + 2);                          // [CTC_WARPS][SP]
    float* occ = vbuf + CTC_WARPS * SP;                               // [CTC_WARPS][C]
    __shared__ float s_nll;
    __shared__ int s_last;

    float* lp = p.lp + (size_t)b * T * C;
    float* la = p.alpha + (size_t)b * T * SP;
    float* lb = p.beta + (size_t)b * T * SP;

    for (int s = tid; s < S; s += CTC_THREADS) {
        int v = p.blank;
        if (s & 1) {
            v = p.targets[(size_t)b * p.tgt_stride + (s >> 1)];
            v = min(max(v, 0), C - 1);
        }
        ext[s] = v;
    }
    // phase 0: log-probs of this utterance, contiguous [T][C]
    for (int t = tid; t < il; t += CTC_THREADS) {
        const float* a = p.act + (int64_t)t * p.st + (int64_t)b * p.sb;
        if (p.is_logits) {
            float m = NSD_NEG_INF;
            for (int c = 0; c < C; ++c) m = fmaxf(m, a[(int64_t)c * p.sc]);
            float sum = 0.f;
            for (int c = 0; c < C; ++c) sum += expf(a[(int64_t)c * p.sc] - m);
            const float lz = logf(sum);
            for (int c = 0; c < C; ++c) lp[(size_t)t * C + c] = a[(int64_t)c * p.sc] - m - lz;
        } else {
            for (int c = 0; c < C; ++c) lp[(size_t)t * C + c] = a[(int64_t)c * p.sc];
        }
    }
    __syncthreads();
    // chain of repeated labels (odd slots): nxt[s] = smallest odd s' > s with ext[s'] == ext[s]
    for (int s = tid; s < S; s += CTC_THREADS) {
        int n = -1;
        if (s & 1)
            for (int s2 = s + 2; s2 < S; s2 += 2)
                if (ext[s2] == ext[s]) { n = s2; break; }
        nxt[s] = n;
        int f = 1;
        if (s & 1)
            for (int s2 = 1; s2 < s; s2 += 2)
                if (ext[s2] == ext[s]) { f = 0; break; }
        isfirst[s] = f;
    }

    // phase 1: warp 0 -> alpha (t ascending), warp 1 -> beta (t descending)
    if (il > 0) {
        float* buf = rowbuf + (warp & 1) * 2 * (SP + 2);
        if (warp == 0) {
            float* prev = buf + 2;                 // two -inf guard cells in front for s-1, s-2
            float* cur = buf + (SP + 2) + 2;
            if (lane == 0) { buf[0] = buf[1] = NSD_NEG_INF; buf[SP + 2] = buf[SP + 3] = NSD_NEG_INF; }
            for (int s = lane; s < S; s += 32) {
                float v = NSD_NEG_INF;
                if (s == 0) v = lp[p.blank];
                else if (s == 1) v = lp[ext[1]];
                prev[s] = v;
                la[s] = v;
            }
            __syncwarp();
            for (int t = 1; t < il; ++t) {
                for (int s = lane; s < S; s += 32) {
                    const int e = ext[s];
                    const float a1 = prev[s], a2 = prev[s - 1];
                    const float a3 = (s > 1 && ext[s - 2] != e) ? prev[s - 2] : NSD_NEG_INF;
                    float v = lse3f(a1, a2, a3);
                    if (v != NSD_NEG_INF) v += lp[(size_t)t * C + e];
                    cur[s] = v;
                    la[(size_t)t * SP + s] = v;
                }
                __syncwarp();
                float* tmp = prev; prev = cur; cur = tmp;
            }
        } else if (warp == 1 && p.grad != nullptr) {
            float* prev = buf;                     // two -inf guard cells after the row for s+1, s+2
            float* cur = buf + (SP + 2);
            for (int s = lane; s < S + 2; s += 32) {
                float v = NSD_NEG_INF;
                if (s == S - 1) v = lp[(size_t)(il - 1) * C + p.blank];
                else if (s == S - 2) v = lp[(size_t)(il - 1) * C + ext[S - 2]];
                prev[s] = v;
                cur[s] = NSD_NEG_INF;
                if (s < S) lb[(size_t)(il - 1) * SP + s] = v;
            }
            __syncwarp();
            for (int t = il - 2; t >= 0; --t) {
                for (int s = lane; s < S; s += 32) {
                    const int e = ext[s];
                    const float b1 = prev[s], b2 = prev[s + 1];
                    const float b3 = (s + 2 < S && ext[s + 2] != e) ? prev[s + 2] : NSD_NEG_INF;
                    float v = lse3f(b1, b2, b3);
                    if (v != NSD_NEG_INF) v += lp[(size_t)t * C + e];
                    cur[s] = v;
                    lb[(size_t)t * SP + s] = v;
                }
                __syncwarp();
                float* tmp = prev; prev = cur; cur = tmp;
            }
        }
    }
    __syncthreads();

    // phase 2: negative log likelihood, then the gradient
    if (tid == 0) {
        float nll;
        if (il == 0) nll = (tl == 0) ? 0.f : INFINITY;
        else {
            const float l1 = la[(size_t)(il - 1) * SP + S - 1];
            const float l2 = (S > 1) ? la[(size_t)(il - 1) * SP + S - 2] : NSD_NEG_INF;
            nll = -lse2f(l1, l2);
        }
        s_nll = nll;
    }
    __syncthreads();
    const float nll = s_nll;
    const bool feasible = (nll != INFINITY) && (nll == nll);      // zero_infinity=True
    if (tid == 0) p.nll[b] = feasible ? nll : 0.f;
    const float scale = p.reduction_mean ? 1.0f / ((float)max(tl, 1) * (float)p.B) : 1.0f;

    float* myv = vbuf + warp * SP;
    float* myocc = occ + warp * C;
    for (int t = warp; t < T && p.grad != nullptr; t += CTC_WARPS) {
        float* g = p.grad + (int64_t)t * p.st + (int64_t)b * p.sb;
        if (t >= il || !feasible) {
            for (int c = lane; c < C; c += 32) g[(int64_t)c * p.sc] = 0.f;
            continue;
        }
        const float* lpt = lp + (size_t)t * C;
        // posterior mass of every lattice slot: exp(alpha + beta - lp + nll)
        float blank_part = 0.f;
        for (int s = lane; s < S; s += 32) {
            const float a = la[(size_t)t * SP + s], bb = lb[(size_t)t * SP + s];
            float v = 0.f;
            if (a != NSD_NEG_INF && bb != NSD_NEG_INF) v = expf(a + bb - lpt[ext[s]] + nll);
            myv[s] = v;
        }
        for (int c = lane; c < C; c += 32) myocc[c] = 0.f;
        __syncwarp();
        // blanks: even slots, lane-strided partial sums in slot order + shuffle tree (fixed order)
        for (int s = 2 * lane; s < S; s += 64) blank_part += myv[s];
        blank_part = warp_sum(blank_part);
        // labels: the lane that owns the first occurrence walks the chain of repeats
        for (int s = 2 * lane + 1; s < S; s += 64) {
            if (isfirst[s]) {
                float acc = 0.f;
                for (int q = s; q >= 0; q = nxt[q]) acc += myv[q];
                myocc[ext[s]] = acc;
            }
        }
        __syncwarp();
        if (lane == 0) myocc[p.blank] += blank_part;
        __syncwarp();
        for (int c = lane; c < C; c += 32) g[(int64_t)c * p.sc] = (expf(lpt[c]) - myocc[c]) * scale;
        __syncwarp();
    }

    // deterministic loss reduction by the last CTA to finish
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int done = atomicAdd(p.counter, 1u);
        s_last = (done == (unsigned int)p.B - 1);
    }
    __syncthreads();
    if (s_last && tid == 0) {
        __threadfence();
        float acc = 0.f;
        for (int i = 0; i < p.B; ++i) {
            const float v = __ldcg(p.nll + i);
            acc += p.reduction_mean ? v / (float)max(min(max(p.tgt_lens[i], 0), (SP - 1) / 2), 1) : v;
        }
        *p.loss = p.reduction_mean ? acc / (float)p.B : acc;
        *p.counter = 0u;
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void greedy_decode_kernel(const float* __restrict__ act, int64_t st, int64_t sb, int64_t sc,
                                     const int32_t* __restrict__ lens, int T, int B, int C, int blank,
                                     int64_t* __restrict__ out, int32_t* __restrict__ out_len) {
    pdl_enter();
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const int L = min(max(lens[b], 0), T);
    int prev = -1, count = 0;
    for (int t0 = 0; t0 < L; t0 += 32) {
        const int t = t0 + lane;
        int id = -1;
        if (t < L) {
            const float* a = act + (int64_t)t * st + (int64_t)b * sb;
            float best = a[0];
            id = 0;
            for (int c = 1; c < C; ++c) {
                const float v = a[(int64_t)c * sc];
                if (v > best || (v != v && best == best)) { best = v; id = c; }   // first max wins; NaN counts as max (torch)
            }
        }
        int left = __shfl_up_sync(0xffffffffu, id, 1);
        if (lane == 0) left = prev;
        const bool keep = (t < L) && (id != left) && (id != blank);
        const unsigned int m = __ballot_sync(0xffffffffu, keep);
        if (keep) out[(size_t)b * T + count + __popc(m & ((1u << lane) - 1u))] = id;
        count += __popc(m);
        const int last_lane = min(31, L - 1 - t0);
        prev = __shfl_sync(0xffffffffu, id, last_lane);
    }
    if (lane == 0) out_len[b] = count;
}

// Row-wise log-softmax, one warp per row (trainer:210, 301).  Same arithmetic as phase 0 of ctc_kernel.
__global__ void log_softmax_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows, int C) {
    pdl_enter();
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* a = in + row * C;
    float m = NSD_NEG_INF;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, a[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(a[c] - m);
    s = warp_sum(s);
    const float lz = logf(s);
    for (int c = lane; c < C; c += 32) out[row * C + c] = a[c] - m - lz;
}

// Levenshtein distance, one warp per utterance, anti-diagonal wavefront over three rolling diagonals.
__global__ void edit_distance_kernel(const int64_t* __restrict__ dec, int dec_stride, const int32_t* __restrict__ dec_len,
                                     const int32_t* __restrict__ tgt, int tgt_stride, const int32_t* __restrict__ tgt_len,
                                     int B, int W, int32_t* __restrict__ ws, int32_t* __restrict__ dist) {
    pdl_enter();
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const int la = tgt_len[b], lb = dec_len[b];           // a = true sequence, b = decoded
    int32_t* d0 = ws + (size_t)b * 3 * W;                  // diagonal k-2, indexed by i (row in a)
    int32_t* d1 = d0 + W;                                  // diagonal k-1
    int32_t* d2 = d1 + W;                                  // diagonal k
    // cell (i,j), i in [0,la], j in [0,lb], k = i+j ; D(i,0)=i ; D(0,j)=j
    for (int k = 0; k <= la + lb; ++k) {
        const int ilo = max(0, k - lb), ihi = min(la, k);
        for (int i = ilo + lane; i <= ihi; i += 32) {
            const int j = k - i;
            int v;
            if (i == 0) v = j;
            else if (j == 0) v = i;
            else {
                const int sub = d0[i - 1] + ((int64_t)tgt[(size_t)b * tgt_stride + i - 1] != dec[(size_t)b * dec_stride + j - 1] ? 1 : 0);
                v = min(min(d1[i - 1] + 1, d1[i] + 1), sub);
            }
            d2[i] = v;
        }
        __syncwarp();
        int32_t* tmp = d0; d0 = d1; d1 = d2; d2 = tmp;
    }
    if (lane == 0) dist[b] = d1[la];
}

}  // namespace nsd

extern "C" {

size_t nsd_ctc_workspace(int T, int B, int C, int max_tgt) {
    const size_t SP = 2 * (size_t)max_tgt + 1;
    return sizeof(float) * ((size_t)B * T * C + 2 * (size_t)B * T * SP) + 256;
}

int nsd_ctc_loss(const float* act, int64_t st, int64_t sb, int64_t sc, int is_logits, const int32_t* targets,
                 int tgt_stride, const int32_t* in_lens, const int32_t* tgt_lens, int T, int B, int C, int blank,
                 int max_tgt, int reduction_mean, float* nll, float* loss, float* grad, void* workspace,
                 size_t workspace_bytes, void* stream) {
    using namespace nsd;
    NSD_CHECK_ARG(T > 0 && B > 0 && C > 0 && max_tgt >= 0, "ctc_loss: bad sizes T=%d B=%d C=%d", T, B, C);
    NSD_CHECK_ARG(blank >= 0 && blank < C, "ctc_loss: blank=%d out of range", blank);
    if (workspace_bytes < nsd_ctc_workspace(T, B, C, max_tgt)) { set_error("ctc_loss: workspace too small"); return NSD_ERR_WORKSPACE; }
    CtcParams p;
    p.act = act; p.st = st; p.sb = sb; p.sc = sc; p.is_logits = is_logits;
    p.targets = targets; p.tgt_stride = tgt_stride; p.in_lens = in_lens; p.tgt_lens = tgt_lens;
    p.T = T; p.B = B; p.C = C; p.blank = blank; p.SP = 2 * max_tgt + 1; p.reduction_mean = reduction_mean;
    p.nll = nll; p.loss = loss; p.grad = grad;
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    p.counter = reinterpret_cast<unsigned int*>(w);
    p.lp = reinterpret_cast<float*>(w + 256);
    p.alpha = p.lp + (size_t)B * T * C;
    p.beta = p.alpha + (size_t)B * T * p.SP;
    cudaStream_t s = (cudaStream_t)stream;
    NSD_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), s));
    const size_t smem = sizeof(int) * 3 * p.SP + sizeof(float) * (4 * (size_t)(p.SP + 2) + CTC_WARPS * (size_t)p.SP + CTC_WARPS * (size_t)C);
    NSD_CHECK_ARG(smem <= 200 * 1024, "ctc_loss: max_tgt=%d C=%d need %zu B shared memory", max_tgt, C, smem);
    if (smem > 48 * 1024) NSD_CUDA(cudaFuncSetAttribute(ctc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nsd::launch_k(ctc_kernel, B, CTC_THREADS, smem, s, p);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}

int nsd_greedy_decode(const float* act, int64_t st, int64_t sb, int64_t sc, const int32_t* lens, int T, int B, int C,
                      int blank, int64_t* out, int32_t* out_len, void* stream) {
    using namespace nsd;
    NSD_CHECK_ARG(T > 0 && B > 0 && C > 0, "greedy_decode: bad sizes");
    nsd::launch_k(greedy_decode_kernel, cdiv(B, 4), 128, 0, (cudaStream_t)stream, act, st, sb, sc, lens, T, B, C, blank, out, out_len);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}

int nsd_log_softmax_f32(const float* in, float* out, int64_t rows, int C, void* stream) {
    using namespace nsd;
    NSD_CHECK_ARG(rows >= 0 && C > 0, "log_softmax: bad sizes");
    if (rows == 0) return NSD_OK;
    nsd::launch_k(log_softmax_kernel, (unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream, in, out, rows, C);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}

size_t nsd_edit_distance_workspace(int B, int max_len) { return sizeof(int32_t) * (size_t)B * 3 * ((size_t)max_len + 1); }

int nsd_edit_distance(const int64_t* dec, int dec_stride, const int32_t* dec_len, const int32_t* tgt, int tgt_stride,
                      const int32_t* tgt_len, int B, int32_t* dist, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace nsd;
    NSD_CHECK_ARG(B > 0, "edit_distance: bad batch");
    const size_t W = workspace_bytes / (sizeof(int32_t) * 3 * (size_t)B);
    NSD_CHECK_ARG(W >= 1, "edit_distance: workspace too small");
    nsd::launch_k(edit_distance_kernel, cdiv(B, 4), 128, 0, (cudaStream_t)stream, dec, dec_stride, dec_len, tgt, tgt_stride, tgt_len, B, (int)W, (int32_t*)workspace, dist);
    NSD_LAUNCH_CHECK();
    return NSD_OK;
}

}  // extern "C"
