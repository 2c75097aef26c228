/*
 * nsd_b200.h -- C ABI of the B200-native GRUDecoder + CTC hot path.
 *
 * The reference (EdwardoSunny/Neural-Speech-Decoder) is pure Python and has no FFI
 * of its own: the boundary its trainer uses is the torch.nn.Module contract
 *   GRUDecoder.forward(neuralInput[B,T,N], dayIdx[B]) -> logits[B,T',C]   src/neural_decoder/model.py:83-123
 *   torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)       src/neural_decoder/neural_decoder_trainer.py:139-141, 213-218
 *   argmax / unique_consecutive / drop-blank greedy decode                src/neural_decoder/neural_decoder_trainer.py:313-320
 * Each entry point below replaces the third-party torch operator(s) that one of
 * those lines reaches on a GPU; the citation names the reference line.  The
 * Python mirror of the reference interface (neural_speech_decoder_b200/model.py,
 * ctc.py) is the only caller.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the comment says "host";
 *   - no torch types, no allocation, no host synchronisation inside any call:
 *     the caller supplies outputs and workspaces and a cudaStream_t (as void*);
 *   - return value: 0 on success, a negative NSD_ERR_* code otherwise;
 *     nsd_last_error() gives a thread-local message for the last failure;
 *   - "time-major" activations are [T', B, F] with row m = t*B + b;
 *   - dtype codes: NSD_F32 = 0, NSD_BF16 = 1.
 */
#ifndef NSD_B200_H
#define NSD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSD_OK 0
#define NSD_ERR_INVALID (-1)     /* bad argument / unsupported shape */
#define NSD_ERR_CUDA (-2)        /* a CUDA runtime call failed */
#define NSD_ERR_WORKSPACE (-3)   /* workspace too small */

#define NSD_F32 0
#define NSD_BF16 1

/* ---- library ---------------------------------------------------------- */
int nsd_version(void);
const char* nsd_last_error(void);
/* number of kernels this library has enqueued in this process (bench.py reports the per-step count). */
unsigned long long nsd_launch_count(void);
/* sm count and compute capability of the current device (host out-pointers). */
int nsd_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- K1: fused front end --------------------------------------------------
 * Replaces F.conv1d(groups=N,"same") + index_select + einsum + Softsign + nn.Unfold
 * (augmentations.py:91; model.py:84-101).
 *   ys[b,t,c]   = sum_k taps[k] * x[b, t-(ntaps-1)/2+k, c]         (zero padded)
 *   pre[b,t,k]  = sum_d ys[b,t,d] * day_w[day[b],d,k] + day_b[day[b],k]
 *   z           = pre / (1+|pre|)
 *   patches[j*B+b, c*K+kk] = z[b, j*S+kk, c]                         (time-major rows)
 * ys and z ([B,T,N] f32) are written for the backward; rows t >= (T'-1)*S+K are
 * left untouched.  err_flag (device-visible int32, may be mapped pinned host
 * memory, may be NULL) is set to 1 if a day index is outside [0,n_days).
 * Fused training augmentation (neural_decoder_trainer.py:194-201): when
 * white_noise_sd / constant_offset_sd are non-zero, x is read as nsd_input_noise(x, ...) with the
 * same seed would have written it -- no extra pass over X.  Distributional parity with the
 * reference's torch.randn (counter-based Philox + Box-Muller here).                        */
int nsd_frontend_fwd(const float* x, const int64_t* day_idx, const float* day_w, const float* day_b,
                     const float* taps, int ntaps, int B, int T, int N, int n_days, int kernel_len,
                     int stride_len, float* ys, float* z, void* patches, int patches_dtype,
                     int* err_flag, float white_noise_sd, float constant_offset_sd, uint64_t noise_seed,
                     void* stream);
/* The augmentation on its own: out[b,t,c] = x[b,t,c] + white_noise_sd * n0[b,t,c] + constant_offset_sd * n1[b,c],
 * n0, n1 standard normal, functions of (seed, element index) only.  x, out: [B,T,N] f32 contiguous (may alias). */
int nsd_input_noise(const float* x, float* out, int B, int T, int N, float white_noise_sd, float constant_offset_sd,
                    uint64_t seed, void* stream);

/* Backward of K1 w.r.t. dayWeights / dayBias (X has no grad): col2im of dpatches,
 * softsign', per-utterance ys^T dpre, then a deterministic segment-reduce over
 * utterances that share a day (autograd of model.py:89-101).
 * workspace: B*N*(N+1) floats.  d_day_w [n_days,N,N], d_day_b [n_days,N] are overwritten. */
int nsd_frontend_bwd(const void* dpatches, int dpatches_dtype, const float* ys, const float* z,
                     const int64_t* day_idx, int B, int T, int N, int n_days, int kernel_len,
                     int stride_len, float* d_day_w, float* d_day_b, void* workspace,
                     size_t workspace_bytes, void* stream);
size_t nsd_frontend_bwd_workspace(int B, int N);

/* ---- K2: GEMMs -----------------------------------------------------------------
 * Row-major C[M,N] = op(A) * op(B) (+ bias[N]) (+ beta*C).  op(A) is [M,K]: stored
 * [M,K] (lda) if !transa, [K,M] (lda) if transa.  op(B) is [K,N]: stored [K,N] (ldb)
 * if !transb, [N,K] (ldb) if transb.  Time-batched W_ih projections, their dgrad /
 * wgrad and the output layer (nn.GRU input GEMMs model.py:50-57,119; nn.Linear
 * model.py:76-81,122).  _f32 is the CUDA-core fp32 path used for fp32 parity;
 * _bf16 is the tcgen05/TMA tensor-core path (bf16 operands, fp32 accumulate in TMEM). */
int nsd_gemm_f32(int transa, int transb, int M, int N, int K, const float* A, int lda, const float* B,
                 int ldb, float* C, int ldc, const float* bias, float beta, void* stream);
int nsd_gemm_bf16(int transa, int transb, int M, int N, int K, const void* A, int lda, const void* B,
                  int ldb, void* C, int ldc, int c_dtype, const float* bias, float beta, void* stream);
/* Two GEMMs of identical shape and layout in ONE launch (operands A0/B0 -> C0 and A1/B1 -> C1, shared leading
 * dimensions; no bias, no beta): the two directions' time-batched W_hh weight gradients each fill only 65 % of the
 * GPU on their own.  Shapes the paired tensor-core form does not take fall back to two nsd_gemm_bf16 calls. */
int nsd_gemm_bf16_x2(int transa, int transb, int M, int N, int K, const void* A0, const void* A1, int lda, const void* B0,
                     const void* B1, int ldb, void* C0, void* C1, int ldc, int c_dtype, void* stream);

/* column sums: out[n] = sum_m a[m*lda + n]  (bias gradients); two fixed-order stages, workspace nsd_colsum_workspace(N). */
int nsd_colsum(const void* a, int a_dtype, int M, int N, int lda, float* out, void* workspace, size_t workspace_bytes,
               void* stream);
size_t nsd_colsum_workspace(int N);
/* dtype conversion of a contiguous buffer, f32 <-> bf16. */
int nsd_cast(const void* src, int src_dtype, void* dst, int dst_dtype, size_t n, void* stream);
/* bf16 copies of a row-major [R,C] matrix (f32 or bf16 source): dst [R,C] (ld_dst) and/or its transpose dstT [C,R]
 * (ld_dstT); either may be NULL.  (nsd_gemm_bf16 takes either operand major natively; the transposed form is what the
 * BPTT kernel's stationary W_hh^T operand is loaded from.) */
int nsd_cast_transpose(const void* src, int src_dtype, int R, int C, int ld_src, void* dst, int ld_dst, void* dstT,
                       int ld_dstT, void* stream);
/* out[B,T,C] <- in[T,B,C] (or the inverse with the roles of T and B swapped by the caller). */
int nsd_swap01_f32(const float* in, float* out, int D0, int D1, int C, void* stream);

/* ---- K3: GRU recurrence ----------------------------------------------------------
 * One layer-direction of nn.GRU with h0 = 0 (model.py:104-119), gate order [r,z,n]:
 *   gh = h_{t-1} W_hh^T + b_hh ; r = s(gi_r+gh_r) ; z = s(gi_z+gh_z)
 *   n = tanh(gi_n + r*gh_n)    ; h_t = (1-z)*n + z*h_{t-1}
 * gi ([T',B,ldgi], already holding x W_ih^T + b_ih) is read at column offset 0 of the
 * pointer given (the caller offsets it for the reverse direction).  hseq is written
 * at the pointer given with row stride ldh.  r,z,n,hn ([T',B,H]) are saved when
 * non-NULL.  reverse != 0 walks t = T'-1..0 over the full padded length.
 * The fp32 path keeps W_hh resident in shared memory across a persistent cooperative
 * launch when it fits, else launches one step at a time.                          */
int nsd_gru_fwd_f32(const float* gi, int ldgi, const float* w_hh, const float* b_hh, int Tp, int B, int H,
                    int reverse, float* hseq, int ldh, float* r, float* z, float* n, float* hn,
                    void* stream);
/* BPTT: dh_t = dhseq_t + carry ; writes dgi[T',B,ldgi] = [dr~,dz~,dn~] and dghn[T',B,H] = dn~*r. */
int nsd_gru_bwd_f32(const float* dhseq, int lddh, const float* hseq, int ldh, const float* r, const float* z,
                    const float* n, const float* hn, const float* w_hh, int Tp, int B, int H, int reverse,
                    float* dgi, int ldgi, float* dghn, void* workspace, size_t workspace_bytes, void* stream);
size_t nsd_gru_bwd_workspace(int B, int H);

/* Tensor-core (bf16 operand, fp32 accumulate/state) form of the same recurrence: ONE cooperative launch walks all
 * timesteps of one layer for D directions at once (direction 1, or direction 0 when reverse0 != 0, runs
 * t = T'-1..0).  w_hh_bf16 is the bf16 copy of [weight_hh_l*, weight_hh_l*_reverse] stacked to [D*3H, H];
 * b_hh is [D*3H]; gi is [T'*B, ldgi] with direction d at column d*3H; hseq (f32) and hseq_bf16 are [T'*B, ldh]
 * with direction d at column d*H (the bf16 copy is what the CTAs exchange between steps and the next layer's GEMM
 * operand); r,z,n,hn are [D][T'*B][H] (all NULL to skip).  Requires H % 64 == 0.  workspace: nsd_gru_tc_workspace.
 * Fused inter-layer dropout (model.py:55): if hdrop_bf16 != NULL it receives nsd_dropout(hseq_bf16, p_drop, seed)
 * (same mask, same rounding) as a second [T'*B, ldh] bf16 tensor -- the next layer's input in train mode.
 * Initial state (streaming inference; D == 1, reverse0 == 0 only): h0 != NULL is [B, ldh] f32 and
 * hseq_bf16 must then hold T'*B + B rows whose FIRST B rows are bf16(h0); the states h_0..h_{T'-1} are written after
 * them.  h0 == NULL is the reference's zero initial state (model.py:104-117). */
int nsd_gru_fwd_bf16(const float* gi, int ldgi, const void* w_hh_bf16, const float* b_hh, int Tp, int B, int H, int D,
                     int reverse0, float* hseq, void* hseq_bf16, int ldh, float* r, float* z, float* n, float* hn,
                     void* hdrop_bf16, float p_drop, uint64_t seed, const float* h0, void* workspace, size_t workspace_bytes,
                     void* stream);
/* BPTT of the above.  w_hhT_bf16 is the bf16 TRANSPOSE of each direction's W_hh stacked to [D*H, 3H].  Writes
 * dgi_bf16 = [dr~,dz~,dn~] and dgh_bf16 = [dr~,dz~,dn~*r], both [T'*B, ldg] bf16 with direction d at column d*3H.
 * p_drop > 0: dhseq is the gradient w.r.t. the DROPPED output (mask of nsd_dropout(., p_drop, seed) over the
 * contiguous [T'*B, lddh] tensor) and is masked/scaled on load.  db_ih/db_hh (both or neither; [D*3H] f32, overwritten):
 * the bias gradients = column sums of dgi / dgh over all T'*B rows, accumulated in fp32 before the bf16 rounding. */
int nsd_gru_bwd_bf16(const float* dhseq, int lddh, const float* hseq, int ldh, const float* r, const float* z,
                     const float* n, const float* hn, const void* w_hhT_bf16, int Tp, int B, int H, int D, int reverse0,
                     void* dgi_bf16, void* dgh_bf16, int ldg, float p_drop, uint64_t seed, float* db_ih, float* db_hh,
                     void* workspace, size_t workspace_bytes, void* stream);
size_t nsd_gru_tc_workspace(int B, int H, int D);

/* inter-layer dropout (nn.GRU dropout=p, train mode, model.py:55): out = x * mask / (1-p),
 * mask from a counter-based generator keyed by (seed, element index); the backward is the
 * same call on the gradient.  Distributional, not bit, parity with torch's generator.   */
int nsd_dropout(const void* x, void* out, int dtype, size_t n, float p, uint64_t seed, void* stream);

/* ---- K4: CTC ----------------------------------------------------------------------
 * torch.nn.CTCLoss(blank, zero_infinity=True) forward AND gradient in one launch
 * (trainer:139-141, 213-218, 242).  `act` is addressed as act[t*st + b*sb + c*sc]; if
 * is_logits != 0 a log-softmax over c is applied first (trainer:210) and `grad` is
 * d/d logits, else `act` are log-probs and `grad` is what torch's ctc backward returns
 * w.r.t. log_probs.  grad uses the same strides as act and is fully overwritten
 * (zeros beyond in_len).  nll[b] is 0 for infeasible utterances.  loss_mean =
 * mean_b nll[b]/max(tgt_len[b],1) and grad is pre-scaled by 1/(max(tgt_len,1)*B) when
 * reduction_mean != 0; otherwise grad is unscaled (d nll[b]) and loss_mean = sum nll.
 * workspace: nsd_ctc_workspace bytes.                                              */
int nsd_ctc_loss(const float* act, int64_t st, int64_t sb, int64_t sc, int is_logits, const int32_t* targets,
                 int tgt_stride, const int32_t* in_lens, const int32_t* tgt_lens, int T, int B, int C,
                 int blank, int max_tgt, int reduction_mean, float* nll, float* loss, float* grad,
                 void* workspace, size_t workspace_bytes, void* stream);
size_t nsd_ctc_workspace(int T, int B, int C, int max_tgt);

/* out[r,:] = log_softmax(in[r,:]) for `rows` contiguous rows of C floats (trainer:210, 301). */
int nsd_log_softmax_f32(const float* in, float* out, int64_t rows, int C, void* stream);

/* ---- K5: greedy decode -----------------------------------------------------------
 * argmax over c (ties -> lowest index) for t < lens[b], collapse repeats, drop blank
 * (trainer:313-320).  out [B,T] int64 (first out_len[b] entries valid).             */
int nsd_greedy_decode(const float* act, int64_t st, int64_t sb, int64_t sc, const int32_t* lens, int T, int B,
                      int C, int blank, int64_t* out, int32_t* out_len, void* stream);
/* Levenshtein distance between decoded and true sequences, one warp per utterance
 * (edit_distance.SequenceMatcher.distance, trainer:322-330).  dist [B] int32.       */
int nsd_edit_distance(const int64_t* dec, int dec_stride, const int32_t* dec_len, const int32_t* tgt,
                      int tgt_stride, const int32_t* tgt_len, int B, int32_t* dist, void* workspace,
                      size_t workspace_bytes, void* stream);
size_t nsd_edit_distance_workspace(int B, int max_len);

/* ---- optimizer step (inside the timed training step) ------------------
 * torch.optim.Adam semantics (trainer:163-169, 259): g = grad*grad_scale + weight_decay*p;
 * m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr/(1-b1^step) * m / (sqrt(v)/sqrt(1-b2^step) + eps).
 * params/grads/exp_avg/exp_avg_sq are HOST arrays of n_tensors DEVICE pointers (f32, contiguous),
 * numel a HOST array.  One launch per 48 tensors.  shadow_bf16 (NULL, or a HOST array whose entries
 * may be NULL): device pointer of a contiguous bf16 copy of the parameter, rewritten with the
 * updated value in the same pass (the tensor-core operand copy the next forward reads).       */
int nsd_adam_step(int n_tensors, void* const* params, const void* const* grads, void* const* exp_avg,
                  void* const* exp_avg_sq, const int64_t* numel, void* const* shadow_bf16, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int step, float grad_scale, void* stream);

/* Streaming inference push for B <= 8 (new: the reference's forward has no state API, model.py:104-119 always starts from
 * h0 = 0 and drops h_n).  ONE cooperative launch consumes S new bins and advances the whole unidirectional decoder by one
 * output frame: smoothing of the S bins that became computable (augmentations.py:91; left/right reach (ntaps-1)/2 and
 * ntaps-1-(ntaps-1)/2), day affine + softsign (model.py:89-93; bf16 operands, fp32 accumulate, like nsd_frontend_fwd's bf16
 * form), slide of the k/s patch (model.py:96-101), per layer l: gi = x_l W_ih^T + b_ih, gh = bf16(h_l) W_hh^T + b_hh, gates
 * r,z,n, h_l <- (1-z) n + z h_l (model.py:50-57, 119), logits = h_{L-1} fc_w^T + fc_b (model.py:122) and
 * ids[b] = argmax_c logits[b,c] (ties -> lowest index, trainer:314).  State carried between calls, all on the device:
 *   rawring f32 [B][ring_rows][N]  raw bins by absolute index mod ring_rows (ring_rows >= ntaps-1+2S);
 *   x0buf bf16 [2][B][N*K]         patch row of frame j in half (j & 1), feature order c*K + k;
 *   n_bins i32 [1]                 bins received so far (the launch adds S);   extra = n_bins - (S*j + K + right) in [0, S)
 *                                  for the newest emitted frame j (constant while every push is S bins);
 *   h f32 [L][B][H], h_bf16 [L][B][H]   state of every layer and its bf16 operand copy, updated in place.
 * bins_in f32 [B][S][N]; w_ih/w_hh/b_ih/b_hh: HOST arrays of L device pointers (bf16 [3H,in_l] / bf16 [3H,H] / f32 [3H] / f32
 * [3H], gate rows r|z|n); logits f32 [B,C]; ids i32 [B] or NULL; err_flag: set to 1 on a day index out of range.
 * H and N*K multiples of 512, N a multiple of 32, S <= 8, L <= 8. */
size_t nsd_stream_push_workspace(int B, int F0, int H, int L);
int nsd_stream_push(const float* bins_in, float* rawring, int ring_rows, const int64_t* day_idx, const float* day_w, const float* day_b,
                    int n_days, const float* taps, int ntaps, void* x0buf_bf16, int* n_bins, int extra, int B, int N, int K, int S, int H,
                    int L, int C, const void* const* w_ih_bf16, const void* const* w_hh_bf16, const void* const* b_ih,
                    const void* const* b_hh, float* h, void* h_bf16, const void* fc_w_bf16, const float* fc_b, float* logits, int* ids,
                    int* err_flag, void* workspace, size_t workspace_bytes, void* stream);

/* Keep n_sms SMs free of the persistent tensor-core GEMM grids from now on (0 = use every SM).  New (no reference
 * counterpart): while parallel.GradSync has a gradient bucket in flight, NCCL's CTAs run on the reserved SMs instead of
 * displacing CTAs of a 148-wide persistent GEMM.  Host-side state, takes effect at the next nsd_gemm_bf16* call. */
int nsd_set_gemm_sm_reserve(int n_sms);

/* n small f32 vectors copied src[i] -> dst[i] (numel[i] elements) in one launch: packs the per-direction nn.GRU bias
 * vectors (model.py:50-57: bias_ih_l{k}[_reverse], bias_hh_l{k}[_reverse]) side by side for the two-direction launches. */
int nsd_multi_copy_f32(int n_tensors, const void* const* src, void* const* dst, const int64_t* numel, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NSD_B200_H */
