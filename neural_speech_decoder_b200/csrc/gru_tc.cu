// K3 tensor-core path: persistent, weight-stationary GRU recurrence (forward and BPTT) on tcgen05.
//
// One cooperative launch walks ALL timesteps of one layer, both directions at once (reference nn.GRU,
// model.py:50-57, 104-119).  A CTA owns U hidden units of one direction: its slice of W_hh (forward: the 3U gate
// rows, K = H; BPTT: U rows of W_hh^T, K = 3H) is TMA-loaded ONCE into shared memory in the K-major 128B-swizzled
// UMMA layout and stays there for the whole sequence.  Every step the CTA streams the previous hidden state
// h_{t-1} [B, H] (BPTT: the previous gate gradients dgh [B, 3H]) -- written to global memory by all CTAs of the
// direction -- through an 8-stage TMA ring as the A operand of UMMA 64 x N x 16 (N = 3U or U), accumulating
// [batch, gate columns] in tensor memory.  In that orientation one TMEM lane = one batch row, so each epilogue
// thread reads r/z/n pre-activations of its own (row, 8 units) with tcgen05.ld and does the gate math without any
// cross-thread exchange.  Steps are separated by a per-direction grid barrier (release/acquire counter in global
// memory) that only the TMA-producer thread waits on; everyone else sleeps on mbarriers.
#include <stdlib.h>

#include "tc_common.cuh"

namespace nsd {
namespace rtc {
using namespace nsd::tc;

constexpr int BT = 64;                  // batch rows per tile (UMMA M = 128 with rows 64..127 unused: TMEM lane = row)
constexpr int STAGES = 8;               // A-operand ring depth
constexpr int A_STAGE = BT * BK * 2;    // 8 KB per k-block
constexpr int CTRL_THREADS = 128;       // warp 0: TMA + grid barrier, warp 1: MMA issue, warp 2: TMEM alloc, warp 3: idle
constexpr int UPT = 8;                  // hidden units per epilogue thread
constexpr int TMEM_COLS = 256;             // 4 accumulation chains x (3U <= 48) columns
constexpr int CHAINS = BK / UMMA_K;
constexpr int CNT_STRIDE = 32;          // uint32 slots between the two directions' step counters (128 B apart)

__device__ __forceinline__ float fast_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return fmaf(0.5f, fast_tanh(0.5f * x), 0.5f); }

struct Smem {
    uint8_t* w;            // stationary weight slice, per k-block [NB rows][128 B]
    uint8_t* a;            // ring [STAGES][A_STAGE]
    uint64_t* full;        // [STAGES]
    uint64_t* empty;       // [STAGES]
    uint64_t* wbar;        // weights landed
    uint64_t* tmem_full;   // MMA -> epilogue
    uint64_t* tmem_empty;  // epilogue -> MMA
    uint32_t* tmem_slot;
};

__device__ __forceinline__ Smem carve(uint8_t* raw, int w_bytes) {
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    Smem s;
    s.w = base;
    s.a = base + w_bytes;                       // w_bytes is a multiple of 1024
    uint64_t* bars = reinterpret_cast<uint64_t*>(s.a + (STAGES + 1) * A_STAGE);   // +1: the M=128 descriptor of the last stage reads 8 KB past it
    s.full = bars; s.empty = bars + STAGES; s.wbar = bars + 2 * STAGES; s.tmem_full = s.wbar + 1; s.tmem_empty = s.wbar + 2;
    s.tmem_slot = reinterpret_cast<uint32_t*>(s.wbar + 3);
    return s;
}

__device__ __forceinline__ void grid_wait(const unsigned int* counter, unsigned int target) {
    if (ld_acquire_u32(counter) >= target) return;
    const long long t0 = clock64();
    unsigned int spins = 0;
    while (ld_acquire_u32(counter) < target) {
        if ((++spins & 0x3FFFu) == 0 && clock64() - t0 > SPIN_CYCLES) {
            printf("nsd gru_tc: grid barrier timeout (block %d, have %u want %u)\n", blockIdx.x, ld_acquire_u32(counter), target);
            __trap();
        }
    }
}

__device__ __forceinline__ void epi_bar_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

struct Common {
    int Tp, B, H, D, U, reverse0, nper;     // nper = CTAs per direction = H / U
    unsigned int* counters;
    int dbg;                                // debug (NSD_GRU_DBG): 1 = skip the MMAs (timing experiment)
    long long* trace;                       // debug (NSD_GRU_TRACE=1): clock64 stamps of block 0, [step < 16][8 events]
};

constexpr int TRACE_STEPS = 16;
__device__ __forceinline__ void stamp(const Common& c, int s, int ev) {
    if (c.trace != nullptr && blockIdx.x == 0 && s < TRACE_STEPS) c.trace[s * 8 + ev] = clock64();
}

// The control warps of both kernels: stream `nkb` k-blocks of A rows [row0, row0+64) per (step, batch tile) through the
// ring and accumulate A * Wslice^T into TMEM.  a_row(step) gives the first A row of the step being consumed.
template <int NB>
__device__ __forceinline__ void control_warps(const Smem& sm, const CUtensorMap* tmA, int warp, int lane, uint32_t tmem_base,
                                              const Common& c, int d, int nkb, int a_col0, bool bptt) {
    const int n_bt = (c.B + BT - 1) / BT;
    const bool rev = (d == 1) || (c.reverse0 != 0);
    // Every CTA of a direction streams the SAME rows; starting each at a different k-block spreads the simultaneous
    // requests over many L2 slices instead of 64 CTAs hammering one 8 KB box at a time.
    const int rot = (int)(((long long)(blockIdx.x - d * c.nper) * nkb) / c.nper);
    if (warp == 0 && lane == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int s = 1; s < c.Tp; ++s) {
            // step s consumes what step s-1 produced: forward h_{t-1}; BPTT dgh of the step handled just before
            int t_src;
            if (!bptt) { const int t = rev ? (c.Tp - 1 - s) : s; t_src = rev ? t + 1 : t - 1; }
            else { const int t = rev ? s : (c.Tp - 1 - s); t_src = rev ? t - 1 : t + 1; }
            grid_wait(c.counters + d * CNT_STRIDE, (unsigned int)(s * c.nper));
            fence_proxy_async();
            stamp(c, s, 0);
            for (int bt = 0; bt < n_bt; ++bt) {
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&sm.empty[stage], phase ^ 1);
                    mbar_expect_tx(&sm.full[stage], A_STAGE);
                    int kk = kb + rot; if (kk >= nkb) kk -= nkb;
                    tma_load_2d(tmA, &sm.full[stage], sm.a + stage * A_STAGE, a_col0 + kk * BK, t_src * c.B + bt * BT);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            stamp(c, s, 1);
        }
    } else if (warp == 1 && lane == 0) {
        // UMMA M = 128 over a 64-row tile: an M = 64 smem-sourced MMA costs ~128 cycles whatever N is; with M = 128 the
        // cost scales with N.  Rows 64..127 alias the next 8 KB of shared memory (finite junk); their TMEM lanes are ignored.
        constexpr uint32_t idesc = make_idesc_bf16(128, NB);
        mbar_wait(sm.wbar, 0);
        int stage = 0; uint32_t phase = 0; uint32_t it = 0;
        for (int s = 1; s < c.Tp; ++s) {
            for (int bt = 0; bt < n_bt; ++bt, ++it) {
                mbar_wait(sm.tmem_empty, (it & 1) ^ 1);
                tcgen05_fence_after();
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&sm.full[stage], phase);
                    tcgen05_fence_after();
                    if (kb == 0 && bt == 0) stamp(c, s, 2);
                    const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(sm.a + stage * A_STAGE));
                    int kk = kb + rot; if (kk >= nkb) kk -= nkb;
                    const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(sm.w + (size_t)kk * NB * 128));
                    // 4 independent accumulation chains (one per 16-wide k-slice of the block), summed by the epilogue:
                    // back-to-back MMAs into ONE small-N accumulator serialise on its read-modify-write latency
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        if (c.dbg != 1) umma_bf16(tmem_base + (uint32_t)(k * NB), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, kb != 0);
                    umma_commit(&sm.empty[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(sm.tmem_full);
                if (bt == 0) stamp(c, s, 3);
            }
        }
    }
}

template <int NB>
__device__ __forceinline__ uint32_t setup(const Smem& sm, int warp, int lane, int n_epi_warps) {
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], 1); }
        mbar_init(sm.wbar, 1); mbar_init(sm.tmem_full, 1); mbar_init(sm.tmem_empty, n_epi_warps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    return *sm.tmem_slot;
}

__device__ __forceinline__ void teardown(int warp, uint32_t tmem_base) {
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8g(const float* p, float (&v)[8]) {   // read-only data (constant for the whole launch)
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8_bf16(__nv_bfloat16* p, const float (&v)[8]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
    *reinterpret_cast<uint4*>(p) = u;
}

// =============================================================================================== forward
struct FwdParams {
    Common c;
    const float* gi; int ldgi;            // [Tp*B, D*3H] = x W_ih^T + b_ih
    const float* b_hh;                    // [D*3H]
    float* hseq; __nv_bfloat16* hseq_bf; int ldh;    // [Tp*B, D*H]; the bf16 copy is what the other CTAs TMA-load
    float* r; float* z; float* n; float* hn;         // [D][Tp*B][H] or null
};

template <int U>
__global__ void __launch_bounds__(CTRL_THREADS + 4 * 32 * (U / UPT), 1)
gru_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH, const FwdParams p) {
    constexpr int NB = 3 * U;
    constexpr int EPI_WARPS = 4 * (U / UPT);
    extern __shared__ uint8_t smem_raw[];
    const Common& c = p.c;
    const int H = c.H, B = c.B;
    const int nkb = H / BK;
    const Smem sm = carve(smem_raw, NB * H * 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = blockIdx.x / c.nper;
    const int u0 = (blockIdx.x - d * c.nper) * U;
    const bool rev = (d == 1) || (c.reverse0 != 0);
    const uint32_t tmem_base = setup<NB>(sm, warp, lane, EPI_WARPS / 2);

    if (warp == 0 && lane == 0) {
        // stationary weights: rows g*H + u0 .. +U of this direction's W_hh, all of K, once
        mbar_expect_tx(sm.wbar, (uint32_t)(NB * H * 2));
        for (int kb = 0; kb < nkb; ++kb)
            for (int g = 0; g < 3; ++g)
                tma_load_2d(&tmW, sm.wbar, sm.w + (size_t)kb * NB * 128 + (size_t)g * U * 128, kb * BK, d * 3 * H + g * H + u0);
    }
    if (warp < 4) {
        control_warps<NB>(sm, &tmH, warp, lane, tmem_base, c, d, nkb, d * H, false);
    } else {
        // ------------------------------------------------------------ epilogue: gates for (row, 8 units)
        const int e = warp - 4, q = e & 3, grp = e >> 2;
        const bool lane_ok = q < 2;                        // TMEM lanes 0..63 hold the 64 batch rows; warps on lanes 64..127 idle
        const int ub = u0 + grp * UPT;                      // first of this thread's 8 units
        const int n_bt = (B + BT - 1) / BT;
        uint32_t it = 0;
        float k_h[8], k_r[8], k_z[8], k_n[8], k_g[8];       // deferred stores (single batch tile)
        bool last_ok = false; size_t last_m = 0;
        for (int s = 0; s < c.Tp; ++s) {
            const int t = rev ? (c.Tp - 1 - s) : s;
            const int tprev = rev ? t + 1 : t - 1;
            for (int bt = 0; bt < n_bt; ++bt) {
                const int b = bt * BT + q * 32 + lane;
                const bool row_ok = lane_ok && b < B;
                const size_t m = (size_t)t * B + b;
                float gi[3][8], hp[8], bn[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { hp[i] = 0.f; bn[i] = 0.f; }
                if (row_ok) {
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        float bh[8];                          // b_hh: L1-resident, re-read instead of pinning 24 registers
                        ld8g(p.gi + m * p.ldgi + d * 3 * H + g * H + ub, gi[g]);
                        ld8g(p.b_hh + d * 3 * H + g * H + ub, bh);
                        if (g < 2) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) gi[g][i] += bh[i];
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) bn[i] = bh[i];
                        }
                    }
                    if (s > 0) {
                        if (n_bt > 1) ld8(p.hseq + ((size_t)tprev * B + b) * p.ldh + d * H + ub, hp);   // written by this very thread
                        else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) hp[i] = k_h[i];                                 // still in registers
                        }
                    }
                }
                float acc[3][8];
                if (s > 0) {
                    mbar_wait(sm.tmem_full, it & 1);
                    tcgen05_fence_after();
                    if (threadIdx.x == CTRL_THREADS && bt == 0) stamp(c, s, 4);
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(grp * UPT);
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        uint32_t raw[CHAINS][8];
#pragma unroll
                        for (int ch = 0; ch < CHAINS; ++ch) tmem_ld_32x8(taddr + (uint32_t)(ch * NB + g * U), raw[ch]);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            acc[g][i] = (__uint_as_float(raw[0][i]) + __uint_as_float(raw[1][i])) + (__uint_as_float(raw[2][i]) + __uint_as_float(raw[3][i]));
                    }
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0 && lane_ok) mbar_arrive(sm.tmem_empty);
                    ++it;
                } else {
#pragma unroll
                    for (int g = 0; g < 3; ++g)
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[g][i] = 0.f;
                }
                if (row_ok) {
                    float rr[8], zz[8], nn[8], gn[8], hv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        rr[i] = fast_sigmoid(gi[0][i] + acc[0][i]);
                        zz[i] = fast_sigmoid(gi[1][i] + acc[1][i]);
                        gn[i] = acc[2][i] + bn[i];
                        nn[i] = fast_tanh(fmaf(rr[i], gn[i], gi[2][i]));
                        hv[i] = fmaf(zz[i], hp[i] - nn[i], nn[i]);          // (1-z)*n + z*h_prev
                    }
                    // the bf16 state is what the other CTAs wait for: store it first, publish, then write the rest
                    st8_bf16(p.hseq_bf + m * p.ldh + d * H + ub, hv);
                    if (n_bt > 1) {
                        st8(p.hseq + m * p.ldh + d * H + ub, hv);
                        if (p.r) {
                            const size_t o = ((size_t)d * c.Tp * B + m) * H + ub;
                            st8(p.r + o, rr); st8(p.z + o, zz); st8(p.n + o, nn); st8(p.hn + o, gn);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { k_h[i] = hv[i]; k_r[i] = rr[i]; k_z[i] = zz[i]; k_n[i] = nn[i]; k_g[i] = gn[i]; }
                    }
                }
                last_ok = row_ok; last_m = m;
            }
            // publish step s: every epilogue thread's stores -> one release increment of the direction's counter
            if (threadIdx.x == CTRL_THREADS) stamp(c, s, 5);
            fence_proxy_async();
            epi_bar_sync(EPI_WARPS * 32);    // idle warps arrive too
            if (threadIdx.x == CTRL_THREADS) {
                stamp(c, s, 6);
                __threadfence();
                atomicAdd(c.counters + d * CNT_STRIDE, 1u);
                stamp(c, s, 7);
            }
            if (n_bt == 1 && last_ok) {                      // off the critical path: nobody else reads these during the launch
                st8(p.hseq + last_m * p.ldh + d * H + ub, k_h);
                if (p.r) {
                    const size_t o = ((size_t)d * c.Tp * B + last_m) * H + ub;
                    st8(p.r + o, k_r); st8(p.z + o, k_z); st8(p.n + o, k_n); st8(p.hn + o, k_g);
                }
            }
        }
    }
    teardown(warp, tmem_base);
}

// =============================================================================================== BPTT
struct BwdParams {
    Common c;
    const float* dhseq; int lddh;         // [Tp*B, D*H] gradient w.r.t. every emitted h_t
    const float* hseq; int ldh;           // forward hidden states (f32)
    const float* r; const float* z; const float* n; const float* hn;   // [D][Tp*B][H]
    __nv_bfloat16* dgi; __nv_bfloat16* dgh; int ldg;    // [Tp*B, D*3H]: [dr~,dz~,dn~] and [dr~,dz~,dn~*r]
    float* carry;                         // [D][B][H]  dh_t * z_t
};

template <int U>
__global__ void __launch_bounds__(CTRL_THREADS + 4 * 32 * (U / UPT), 1)
gru_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmWT, const __grid_constant__ CUtensorMap tmG, const BwdParams p) {
    constexpr int NB = U;
    constexpr int EPI_WARPS = 4 * (U / UPT);
    extern __shared__ uint8_t smem_raw[];
    const Common& c = p.c;
    const int H = c.H, B = c.B;
    const int nkb = 3 * H / BK;
    const Smem sm = carve(smem_raw, NB * 3 * H * 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = blockIdx.x / c.nper;
    const int u0 = (blockIdx.x - d * c.nper) * U;
    const bool rev = (d == 1) || (c.reverse0 != 0);
    const uint32_t tmem_base = setup<NB>(sm, warp, lane, EPI_WARPS / 2);

    if (warp == 0 && lane == 0) {
        // stationary weights: rows u0 .. u0+U of this direction's W_hh^T [H, 3H], all of K = 3H, once
        mbar_expect_tx(sm.wbar, (uint32_t)(NB * 3 * H * 2));
        for (int kb = 0; kb < nkb; ++kb)
            tma_load_2d(&tmWT, sm.wbar, sm.w + (size_t)kb * NB * 128, kb * BK, d * H + u0);
    }
    if (warp < 4) {
        control_warps<NB>(sm, &tmG, warp, lane, tmem_base, c, d, nkb, d * 3 * H, true);
    } else {
        const int e = warp - 4, q = e & 3, grp = e >> 2;
        const bool lane_ok = q < 2;                        // TMEM lanes 0..63 hold the 64 batch rows; warps on lanes 64..127 idle
        const int ub = u0 + grp * UPT;
        const int n_bt = (B + BT - 1) / BT;
        uint32_t it = 0;
        float k_r[8], k_z[8], k_n[8], k_c[8];               // deferred stores / carry kept in registers (single batch tile)
        bool last_ok = false; size_t last_m = 0;
        for (int s = 0; s < c.Tp; ++s) {
            const int t = rev ? s : (c.Tp - 1 - s);                  // BPTT visits time in the opposite order of the forward pass
            const int tprev = rev ? t + 1 : t - 1;                   // forward-time predecessor (source of h_{t-1})
            const bool has_prev = rev ? (t + 1 < c.Tp) : (t > 0);
            for (int bt = 0; bt < n_bt; ++bt) {
                const int b = bt * BT + q * 32 + lane;
                const bool row_ok = lane_ok && b < B;
                const size_t m = (size_t)t * B + b;
                float dh[8], rr[8], zz[8], nn[8], gn[8], hp[8], cr[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { hp[i] = 0.f; cr[i] = 0.f; }
                if (row_ok) {
                    const size_t o = ((size_t)d * c.Tp * B + m) * H + ub;
                    ld8g(p.dhseq + m * p.lddh + d * H + ub, dh);
                    ld8g(p.r + o, rr); ld8g(p.z + o, zz); ld8g(p.n + o, nn); ld8g(p.hn + o, gn);
                    if (has_prev) ld8g(p.hseq + ((size_t)tprev * B + b) * p.ldh + d * H + ub, hp);
                    if (s > 0) {
                        if (n_bt > 1) ld8(p.carry + ((size_t)d * B + b) * H + ub, cr);          // written by this very thread
                        else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) cr[i] = k_c[i];                         // still in registers
                        }
                    }
                }
                float acc[8];
                if (s > 0) {
                    mbar_wait(sm.tmem_full, it & 1);
                    tcgen05_fence_after();
                    if (threadIdx.x == CTRL_THREADS && bt == 0) stamp(c, s, 4);
                    uint32_t raw[CHAINS][8];
#pragma unroll
                    for (int ch = 0; ch < CHAINS; ++ch) tmem_ld_32x8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * NB + grp * UPT), raw[ch]);
                    tmem_ld_wait();
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0 && lane_ok) mbar_arrive(sm.tmem_empty);
                    ++it;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        acc[i] = (__uint_as_float(raw[0][i]) + __uint_as_float(raw[1][i])) + (__uint_as_float(raw[2][i]) + __uint_as_float(raw[3][i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
                }
                if (row_ok) {
                    float drt[8], dzt[8], dnt[8], dgn[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float dht = dh[i] + cr[i] + acc[i];
                        const float dn = dht * (1.0f - zz[i]);
                        const float dz = dht * (hp[i] - nn[i]);
                        dnt[i] = dn * (1.0f - nn[i] * nn[i]);
                        dzt[i] = dz * zz[i] * (1.0f - zz[i]);
                        drt[i] = dnt[i] * gn[i] * rr[i] * (1.0f - rr[i]);
                        dgn[i] = dnt[i] * rr[i];
                        cr[i] = dht * zz[i];
                    }
                    __nv_bfloat16* gi_row = p.dgi + m * p.ldg + d * 3 * H + ub;
                    __nv_bfloat16* gh_row = p.dgh + m * p.ldg + d * 3 * H + ub;
                    st8_bf16(gh_row, drt); st8_bf16(gh_row + H, dzt); st8_bf16(gh_row + 2 * H, dgn);   // what the other CTAs wait for
                    if (n_bt > 1) {
                        st8_bf16(gi_row, drt); st8_bf16(gi_row + H, dzt); st8_bf16(gi_row + 2 * H, dnt);
                        st8(p.carry + ((size_t)d * B + b) * H + ub, cr);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { k_r[i] = drt[i]; k_z[i] = dzt[i]; k_n[i] = dnt[i]; k_c[i] = cr[i]; }
                    }
                }
                last_ok = row_ok; last_m = m;
            }
            if (threadIdx.x == CTRL_THREADS) stamp(c, s, 5);
            fence_proxy_async();
            epi_bar_sync(EPI_WARPS * 32);    // idle warps arrive too
            if (threadIdx.x == CTRL_THREADS) {
                stamp(c, s, 6);
                __threadfence();
                atomicAdd(c.counters + d * CNT_STRIDE, 1u);
                stamp(c, s, 7);
            }
            if (n_bt == 1 && last_ok) {                      // off the critical path
                __nv_bfloat16* gi_row = p.dgi + last_m * p.ldg + d * 3 * H + ub;
                st8_bf16(gi_row, k_r); st8_bf16(gi_row + H, k_z); st8_bf16(gi_row + 2 * H, k_n);
            }
        }
    }
    teardown(warp, tmem_base);
}

// ---------------------------------------------------------------- host side
static int pick_units(int H, int D) {
    // as many CTAs as can be co-resident (one per SM): 8 units per CTA if that fits, else 16
    if ((H % 8) == 0 && D * (H / 8) <= sm_count()) return 8;
    if ((H % 16) == 0 && D * (H / 16) <= sm_count()) return 16;
    return 0;
}

static size_t smem_bytes(int w_bytes) { return (size_t)w_bytes + (STAGES + 1) * A_STAGE + 1024 + 256; }

template <typename Kern, typename P>
static int launch_coop(Kern kern, int grid, int threads, size_t smem, const CUtensorMap& m0, const CUtensorMap& m1, const P& p, cudaStream_t s) {
    NSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    NSD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if ((long long)per_sm * sm_count() < grid) { set_error("gru_tc: %d CTAs cannot be co-resident (%d per SM)", grid, per_sm); return NSD_ERR_INVALID; }
    void* args[] = {(void*)&m0, (void*)&m1, (void*)&p};
    NSD_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(threads), args, smem, s));
    count_launch(1);
    return NSD_OK;
}

// Debug aid: NSD_GRU_TRACE=1 prints block 0's per-step event times (SM cycles relative to the step's barrier pass).
static long long* trace_begin() {
    const char* e = getenv("NSD_GRU_TRACE");
    if (!e || e[0] != '1') return nullptr;
    long long* d = nullptr;
    if (cudaMalloc(&d, sizeof(long long) * TRACE_STEPS * 8) != cudaSuccess) return nullptr;
    cudaMemset(d, 0, sizeof(long long) * TRACE_STEPS * 8);
    return d;
}
static void trace_end(const char* who, long long* d, cudaStream_t s) {
    if (!d) return;
    long long h[TRACE_STEPS * 8];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    fprintf(stderr, "[%s trace, cycles] step: barrier->tma_issued mma_first_full mma_committed epi_wake epi_stored epi_bar published | step period\n", who);
    for (int st = 1; st < TRACE_STEPS; ++st) {
        const long long* r = h + st * 8;
        if (r[0] == 0) break;
        fprintf(stderr, "  s=%2d: %6lld %6lld %6lld %6lld %6lld %6lld %6lld | %6lld\n", st, r[1] - r[0], r[2] - r[0], r[3] - r[0], r[4] - r[0],
                r[5] - r[0], r[6] - r[0], r[7] - r[0], st > 1 ? r[0] - h[(st - 1) * 8] : 0LL);
    }
}

static int check_shape(const char* who, int Tp, int B, int H, int D, int* U) {
    if (!(Tp > 0 && B > 0 && H > 0 && (D == 1 || D == 2))) { set_error("%s: bad sizes Tp=%d B=%d H=%d D=%d", who, Tp, B, H, D); return NSD_ERR_INVALID; }
    if (H % BK != 0) { set_error("%s: hidden size %d must be a multiple of 64 on the tensor-core path", who, H); return NSD_ERR_INVALID; }
    *U = pick_units(H, D);
    if (*U == 0 || (size_t)3 * (*U) * H * 2 > 160 * 1024) { set_error("%s: hidden size %d x %d directions does not fit weight-stationary on this GPU", who, H, D); return NSD_ERR_INVALID; }
    return NSD_OK;
}

}  // namespace rtc
}  // namespace nsd

extern "C" {

size_t nsd_gru_tc_workspace(int B, int H, int D) { return 256 + sizeof(float) * (size_t)D * B * H; }

int nsd_gru_fwd_bf16(const float* gi, int ldgi, const void* w_hh_bf16, const float* b_hh, int Tp, int B, int H, int D,
                     int reverse0, float* hseq, void* hseq_bf16, int ldh, float* r, float* z, float* n, float* hn,
                     void* workspace, size_t workspace_bytes, void* stream) {
    using namespace nsd;
    using namespace nsd::rtc;
    int U = 0;
    int rc = check_shape("gru_fwd_bf16", Tp, B, H, D, &U);
    if (rc) return rc;
    NSD_CHECK_ARG((r && z && n && hn) || (!r && !z && !n && !hn), "gru_fwd_bf16: save pointers must be all set or all NULL");
    NSD_CHECK_ARG((ldgi % 4) == 0 && (ldh % 8) == 0, "gru_fwd_bf16: ldgi must be a multiple of 4 and ldh of 8");
    if (workspace_bytes < nsd_gru_tc_workspace(B, H, D)) { set_error("gru_fwd_bf16: workspace too small"); return NSD_ERR_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    NSD_CUDA(cudaMemsetAsync(workspace, 0, 256, s));
    CUtensorMap tmW, tmH;
    rc = make_bf16_map(&tmW, w_hh_bf16, (long long)D * 3 * H, H, H, U);
    if (rc) return rc;
    rc = make_bf16_map(&tmH, hseq_bf16, (long long)Tp * B, D * H, ldh, BT);
    if (rc) return rc;
    FwdParams p;
    long long* tr = trace_begin();
    p.c = {Tp, B, H, D, U, reverse0, H / U, reinterpret_cast<unsigned int*>(workspace), getenv("NSD_GRU_DBG") ? atoi(getenv("NSD_GRU_DBG")) : 0, tr};
    p.gi = gi; p.ldgi = ldgi; p.b_hh = b_hh; p.hseq = hseq; p.hseq_bf = reinterpret_cast<__nv_bfloat16*>(hseq_bf16); p.ldh = ldh;
    p.r = r; p.z = z; p.n = n; p.hn = hn;
    const int grid = D * (H / U);
    const size_t smem = smem_bytes(3 * U * H * 2);
    rc = (U == 8) ? launch_coop(gru_fwd_tc_kernel<8>, grid, CTRL_THREADS + 128, smem, tmW, tmH, p, s)
                  : launch_coop(gru_fwd_tc_kernel<16>, grid, CTRL_THREADS + 256, smem, tmW, tmH, p, s);
    trace_end("gru_fwd_bf16", tr, s);
    return rc;
}

int nsd_gru_bwd_bf16(const float* dhseq, int lddh, const float* hseq, int ldh, const float* r, const float* z,
                     const float* n, const float* hn, const void* w_hhT_bf16, int Tp, int B, int H, int D, int reverse0,
                     void* dgi_bf16, void* dgh_bf16, int ldg, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace nsd;
    using namespace nsd::rtc;
    int U = 0;
    int rc = check_shape("gru_bwd_bf16", Tp, B, H, D, &U);
    if (rc) return rc;
    NSD_CHECK_ARG((lddh % 4) == 0 && (ldh % 4) == 0 && (ldg % 8) == 0, "gru_bwd_bf16: leading dimensions must be multiples of 4 (f32) / 8 (bf16)");
    if (workspace_bytes < nsd_gru_tc_workspace(B, H, D)) { set_error("gru_bwd_bf16: workspace too small"); return NSD_ERR_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    NSD_CUDA(cudaMemsetAsync(workspace, 0, 256, s));
    CUtensorMap tmWT, tmG;
    rc = make_bf16_map(&tmWT, w_hhT_bf16, (long long)D * H, 3 * H, 3 * H, U);
    if (rc) return rc;
    rc = make_bf16_map(&tmG, dgh_bf16, (long long)Tp * B, D * 3 * H, ldg, BT);
    if (rc) return rc;
    BwdParams p;
    long long* tr = trace_begin();
    p.c = {Tp, B, H, D, U, reverse0, H / U, reinterpret_cast<unsigned int*>(workspace), getenv("NSD_GRU_DBG") ? atoi(getenv("NSD_GRU_DBG")) : 0, tr};
    p.dhseq = dhseq; p.lddh = lddh; p.hseq = hseq; p.ldh = ldh; p.r = r; p.z = z; p.n = n; p.hn = hn;
    p.dgi = reinterpret_cast<__nv_bfloat16*>(dgi_bf16); p.dgh = reinterpret_cast<__nv_bfloat16*>(dgh_bf16); p.ldg = ldg;
    p.carry = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + 256);
    const int grid = D * (H / U);
    const size_t smem = smem_bytes(U * 3 * H * 2);
    rc = (U == 8) ? launch_coop(gru_bwd_tc_kernel<8>, grid, CTRL_THREADS + 128, smem, tmWT, tmG, p, s)
                  : launch_coop(gru_bwd_tc_kernel<16>, grid, CTRL_THREADS + 256, smem, tmWT, tmG, p, s);
    trace_end("gru_bwd_bf16", tr, s);
    return rc;
}

}  // extern "C"
