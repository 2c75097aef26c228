// K3 tensor-core path: persistent, weight-stationary GRU recurrence (forward and BPTT) on tcgen05.
//
// One cooperative launch walks ALL timesteps of one layer, both directions at once (reference nn.GRU,
// model.py:50-57, 104-119).  Measured on B200 (tests/trace_gru.py, profiles/): a shared-memory-sourced tcgen05.mma
// costs ~128 cycles per K=16 slab whatever N (<= 256) is, and an SM can ingest ~28 B/cycle when every SM pulls
// from L2 at once.  So the work is cut along K, not along the output columns:
//
//   * a cluster of CS = 4 CTAs owns 64 hidden units of one direction; CTA j of the cluster holds, stationary in
//     shared memory for the whole sequence (TMA-loaded once, K-major 128B-swizzled UMMA layout), the slice of W_hh
//     that multiplies ITS QUARTER of the reduction dimension for ALL of the cluster's output columns
//     (forward: 3 gates x 64 units = N 192, K = H/4;  BPTT: W_hh^T, N = 64 units, K = 3H/4);
//   * every step it TMA-loads only its quarter of the previous state (h_{t-1} or dgh, [64 rows, K/4]; all boxes in
//     flight at once, no ring), issues K/64 tcgen05.mma (UMMA M=128 over the 64-row tile, fp32 accumulators in TMEM),
//     and the four partial sums meet through distributed shared memory: each epilogue thread reads its TMEM lane
//     (= one batch row) and pushes the columns owned by the three peers straight into their inboxes with
//     st.async (mbarrier complete_tx signalling), keeps its own 16 units, adds the three partials it receives and
//     does the gate math for (row, 8 units) in registers -- no cross-thread exchange inside the CTA;
//   * steps are separated by a per-direction grid barrier (release/acquire counter in global memory) that only the
//     TMA-producer thread polls; the state the other CTAs need (bf16 h / dgh) is stored and published first, the
//     fp32 state and the saved activations are written after the publish, off the critical path.
#include <stdlib.h>

#include "tc_common.cuh"

namespace nsd {
namespace rtc {
using namespace nsd::tc;

constexpr int BT = 64;                  // batch rows per tile (UMMA M = 128 with rows 64..127 unused: TMEM lane = row)
constexpr int CS = 4;                   // cluster size = K split
constexpr int U = 16;                   // hidden units owned by one CTA (gate math)
constexpr int UC = U * CS;              // hidden units owned by one cluster (UMMA N = 3*UC forward, UC in BPTT)
constexpr int A_BOX = BT * BK * 2;      // 8 KB: one [64 rows x 64 k] bf16 box
constexpr int CTRL_THREADS = 128;       // warp 0: TMA + grid barrier, warp 1: MMA issue, warp 2: TMEM alloc, warp 3: idle
constexpr int EPI_WARPS = 8;            // warps 4..11; TMEM lanes 0..63 are reachable from the warps with (warp & 3) < 2
constexpr int THREADS = CTRL_THREADS + EPI_WARPS * 32;
constexpr int UPT = 8;                  // hidden units per epilogue thread
constexpr int TMEM_COLS = 256;
constexpr int CNT_STRIDE = 32;          // uint32 slots between the two directions' step counters (128 B apart)
constexpr int TRACE_STEPS = 16;

__device__ __forceinline__ float fast_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return fmaf(0.5f, fast_tanh(0.5f * x), 0.5f); }

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    unsigned int spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if ((++spins & 0xFFFFu) == 0 && clock64() - t0 > SPIN_CYCLES) {
            printf("nsd gru_tc: inbox timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
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
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory"); }

struct Common {
    int Tp, B, H, D, reverse0, nper;        // nper = CTAs per direction = H / U
    unsigned int* counters;
    int dbg;                                // debug (NSD_GRU_DBG): 1 = skip the MMAs (timing experiment)
    long long* trace;                       // debug (NSD_GRU_TRACE=1): clock64 stamps of block 0, [step < 16][8 events]
};
__device__ __forceinline__ void stamp(const Common& c, int s, int ev) {
    if (c.trace == nullptr) return;
    if (blockIdx.x == 0 && s < TRACE_STEPS) c.trace[s * 8 + ev] = clock64();
    if (s == 8) {                           // every block, one step, global nanosecond timer: skew across CTAs
        unsigned long long g;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        c.trace[TRACE_STEPS * 8 + blockIdx.x * 8 + ev] = (long long)g;
    }
}

// Shared memory: [A boxes: nkb x 8 KB][W slice: nkb x NB x 128 B][inbox: CS x NIN floats x 64 rows][barriers]
// (A first: the M=128 descriptor of the last box reads 8 KB past it, i.e. into the weights -- finite junk, ignored lanes)
struct Smem {
    uint8_t* a; uint8_t* w; __nv_bfloat16* inbox; __nv_bfloat16* outbox;   // (CS-1) slots each: [slot][gate][row][16 bf16]
    uint64_t* full;        // [nkb] A box landed
    uint64_t* wbar; uint64_t* tmem_full; uint64_t* tmem_empty; uint64_t* inbox_bar;
    uint32_t* tmem_slot;
};
// nkb_max (the same in every CTA of the launch) fixes the layout, so that a peer's inbox sits at the same offset as mine
__device__ __forceinline__ Smem carve(uint8_t* raw, int nkb_max, int nb, int inbox_bytes) {
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    Smem s;
    s.a = base;
    s.w = base + (size_t)nkb_max * A_BOX;
    const size_t w_bytes = (size_t)nkb_max * nb * 128;
    s.inbox = reinterpret_cast<__nv_bfloat16*>(s.w + (w_bytes < (size_t)A_BOX ? (size_t)A_BOX : w_bytes));
    s.outbox = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(s.inbox) + inbox_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s.outbox) + inbox_bytes);
    s.full = bars; s.wbar = bars + nkb_max; s.tmem_full = s.wbar + 1; s.tmem_empty = s.wbar + 2; s.inbox_bar = s.wbar + 3;
    s.tmem_slot = reinterpret_cast<uint32_t*>(s.wbar + 4);
    return s;
}
static size_t smem_bytes(int nkb_max, int nb, int inbox_bytes) {
    const size_t w_bytes = (size_t)nkb_max * nb * 128;
    return (size_t)nkb_max * A_BOX + (w_bytes < (size_t)A_BOX ? (size_t)A_BOX : w_bytes) + 2 * (size_t)inbox_bytes + (nkb_max + 8) * 8 + 1024 + 64;
}

__device__ __forceinline__ uint32_t setup(const Smem& sm, int warp, int lane, int nkb_max) {
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < nkb_max; ++i) mbar_init(&sm.full[i], 1);
        mbar_init(sm.wbar, 1); mbar_init(sm.tmem_full, 1); mbar_init(sm.tmem_empty, EPI_WARPS / 2); mbar_init(sm.inbox_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    __syncwarp();
    cluster_sync_all();                 // every CTA's barriers exist before any peer pushes into its inbox
    tcgen05_fence_after();
    return *sm.tmem_slot;
}
__device__ __forceinline__ void teardown(int warp, uint32_t tmem_base) {
    tcgen05_fence_before();
    __syncthreads();
    __syncwarp();
    cluster_sync_all();                 // nobody leaves while a peer may still push into its shared memory
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// Control warps of both kernels.  Step s >= 1 consumes the rows the direction produced in step s-1.
// A CTA whose share of K is empty (tiny H) issues nothing and contributes zero partial sums.
template <int NB>
__device__ __forceinline__ void control_warps(const Smem& sm, const CUtensorMap* tmA, int warp, int lane, uint32_t tmem_base,
                                              const Common& c, int d, int kb_lo, int nkb, int a_col0, bool bptt) {
    const int n_bt = (c.B + BT - 1) / BT;
    const bool rev = (d == 1) || (c.reverse0 != 0);
    if (nkb == 0) return;
    // Batch tiles of 64 rows are independent sequences: the whole time loop runs per tile (weights stay resident).
    if (warp == 0 && lane == 0) {
        for (int bt = 0; bt < n_bt; ++bt) {
            for (int s = 1; s < c.Tp; ++s) {
                int t_src;
                if (!bptt) { const int t = rev ? (c.Tp - 1 - s) : s; t_src = rev ? t + 1 : t - 1; }
                else { const int t = rev ? s : (c.Tp - 1 - s); t_src = rev ? t - 1 : t + 1; }
                grid_wait(c.counters + d * CNT_STRIDE, (unsigned int)((bt * c.Tp + s) * c.nper));
                fence_proxy_async();
                if (bt == 0) stamp(c, s, 0);
                for (int kb = 0; kb < nkb; ++kb) {            // all boxes in flight at once; last step's MMAs retired long ago
                    mbar_expect_tx(&sm.full[kb], A_BOX);
                    tma_load_2d(tmA, &sm.full[kb], sm.a + kb * A_BOX, a_col0 + (kb_lo + kb) * BK, t_src * c.B + bt * BT);
                }
                if (bt == 0) stamp(c, s, 1);
            }
        }
    } else if (warp == 1 && lane == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, NB);
        mbar_wait(sm.wbar, 0);
        uint32_t it = 0;
        for (int bt = 0; bt < n_bt; ++bt) {
            for (int s = 1; s < c.Tp; ++s, ++it) {
                mbar_wait(sm.tmem_empty, (it & 1) ^ 1);
                tcgen05_fence_after();
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&sm.full[kb], it & 1);
                    tcgen05_fence_after();
                    if (kb == 0 && bt == 0) stamp(c, s, 2);
                    const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(sm.a + kb * A_BOX));
                    const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(sm.w + (size_t)kb * NB * 128));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        if (c.dbg != 1) umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                }
                umma_commit(sm.tmem_full);
                if (bt == 0) stamp(c, s, 3);
            }
        }
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

// Exchange of one gate block: this thread's TMEM lane holds, at columns col0 + 16*p + 8*grp .. +8, the partial sums of
// the 8 units (grp) that cluster rank p owns.  Keep p == me; stage the rest in the local outbox slot of peer p
// ([slot][gate][row][16 floats], slot = (p - me - 1) mod CS).  `have` = this CTA ran MMAs this step (nkb > 0).
template <int NG>
__device__ __forceinline__ void stage_gate(uint32_t trow, int col0, int gate, int me, int row, int grp, const Smem& sm,
                                           bool have, float (&own)[8]) {
    uint32_t raw[CS][8];
    if (have) {
#pragma unroll
        for (int p = 0; p < CS; ++p) tmem_ld_32x8(trow + (uint32_t)(col0 + U * p + UPT * grp), raw[p]);
        tmem_ld_wait();
    } else {
#pragma unroll
        for (int p = 0; p < CS; ++p)
#pragma unroll
            for (int i = 0; i < 8; ++i) raw[p][i] = 0u;
    }
#pragma unroll
    for (int p = 0; p < CS; ++p) {
        if (p == me) {
#pragma unroll
            for (int i = 0; i < 8; ++i) own[i] = __uint_as_float(raw[p][i]);
        } else {
            const int slot = (p - me - 1 + CS) % CS;
            // partial sums travel as bf16 (the DSMEM link moves ~10 B/cycle/SM): 3 of the 4 addends carry 2^-9 relative rounding
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(raw[p][i]);
            st8_bf16(sm.outbox + ((size_t)(slot * NG + gate) * BT + row) * U + UPT * grp, v);
        }
    }
}
// One bulk shared->distributed-shared copy per peer: my outbox slot for peer p lands in p's inbox slot for me
// ((me - p - 1) mod CS) and completes p's inbox mbarrier with the byte count.  Issued by one thread after the staging
// threads have fenced (generic -> async proxy) and synchronised.
template <int NG>
__device__ __forceinline__ void send_outbox(const Smem& sm, int me) {
    constexpr uint32_t BYTES = NG * BT * U * 2;
#pragma unroll
    for (int p = 0; p < CS; ++p) {
        if (p == me) continue;
        const int out_slot = (p - me - 1 + CS) % CS, in_slot = (me - p - 1 + CS) % CS;
        const uint32_t src = smem_u32(sm.outbox + (size_t)out_slot * NG * BT * U);
        const uint32_t dst = map_to_cta(smem_u32(sm.inbox + (size_t)in_slot * NG * BT * U), (uint32_t)p);
        const uint32_t bar = map_to_cta(smem_u32(sm.inbox_bar), (uint32_t)p);
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "r"(src), "r"(BYTES), "r"(bar) : "memory");
    }
}
__device__ __forceinline__ void useful_bar_sync() { asm volatile("bar.sync 2, %0;" ::"n"(EPI_WARPS * 16) : "memory"); }   // the 4 warps on TMEM lanes 0..63
template <int NG>
__device__ __forceinline__ void add_inbox(const Smem& sm, int gate, int row, int grp, float (&acc)[8]) {
#pragma unroll
    for (int slot = 0; slot < CS - 1; ++slot) {
        const uint4 u = *reinterpret_cast<const uint4*>(sm.inbox + ((size_t)(slot * NG + gate) * BT + row) * U + UPT * grp);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {                        // bf16 -> f32 is a 16-bit shift
            acc[2 * i] += __uint_as_float(w[i] << 16);
            acc[2 * i + 1] += __uint_as_float(w[i] & 0xFFFF0000u);
        }
    }
}

// =============================================================================================== forward
struct FwdParams {
    Common c;
    const float* gi; int ldgi;            // [Tp*B, D*3H] = x W_ih^T + b_ih
    const float* b_hh;                    // [D*3H]
    float* hseq; __nv_bfloat16* hseq_bf; int ldh;    // [Tp*B, D*H]; the bf16 copy is what the other CTAs TMA-load
    float* r; float* z; float* n; float* hn;         // [D][Tp*B][H] or null
};

__global__ void __launch_bounds__(THREADS, 1)
gru_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH, const FwdParams p) {
    constexpr int NB = 3 * UC;                       // 192 gate columns of the cluster
    constexpr int INBOX = (CS - 1) * 3 * BT * U * 2;
    extern __shared__ uint8_t smem_raw[];
    const Common& c = p.c;
    const int H = c.H, B = c.B;
    const int me = (int)cluster_rank();
    const int nkb_all = H / BK, nkb_max = (nkb_all + CS - 1) / CS;
    const int kb_lo = me * nkb_all / CS, nkb = (me + 1) * nkb_all / CS - kb_lo;     // this CTA's quarter of K
    const Smem sm = carve(smem_raw, nkb_max, NB, INBOX);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = blockIdx.x / c.nper;
    const int uc0 = ((blockIdx.x - d * c.nper) / CS) * UC;          // first unit of the cluster
    const bool rev = (d == 1) || (c.reverse0 != 0);
    const uint32_t tmem_base = setup(sm, warp, lane, nkb_max);

    if (warp == 0 && lane == 0 && nkb > 0) {
        // stationary weights: for each of the CTA's k-blocks, rows [r | z | n] x 64 cluster units of this direction's W_hh
        mbar_expect_tx(sm.wbar, (uint32_t)(nkb * NB * 128));
        for (int kb = 0; kb < nkb; ++kb)
            for (int g = 0; g < 3; ++g)
                tma_load_2d(&tmW, sm.wbar, sm.w + (size_t)kb * NB * 128 + (size_t)g * UC * 128, (kb_lo + kb) * BK, d * 3 * H + g * H + uc0);
    }
    if (warp < 4) {
        control_warps<NB>(sm, &tmH, warp, lane, tmem_base, c, d, kb_lo, nkb, d * H, false);
    } else {
        // ------------------------------------------------------------ epilogue: gates for (row, 8 units)
        const int e = warp - 4, q = e & 3, grp = e >> 2;
        const bool lane_ok = q < 2;                        // TMEM lanes 0..63 hold the 64 batch rows; warps on lanes 64..127 idle
        const int row = q * 32 + lane;
        const int ub = uc0 + me * U + grp * UPT;           // first of this thread's 8 units
        const int n_bt = (B + BT - 1) / BT;
        uint32_t it = 0;
        for (int bt = 0; bt < n_bt; ++bt) {
            const int b = bt * BT + row;
            const bool row_ok = lane_ok && b < B;
            float k_h[8];                                    // h_{t-1} of this thread's (row, 8 units), fp32, in registers
#pragma unroll
            for (int i = 0; i < 8; ++i) k_h[i] = 0.f;
            for (int s = 0; s < c.Tp; ++s) {
                const int t = rev ? (c.Tp - 1 - s) : s;
                const size_t m = (size_t)t * B + b;
                float gi[3][8], bn[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) bn[i] = 0.f;
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
                }
                float acc[3][8];
#pragma unroll
                for (int g = 0; g < 3; ++g)
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[g][i] = 0.f;
                if (s > 0) {
                    if (threadIdx.x == CTRL_THREADS) mbar_expect_tx(sm.inbox_bar, (uint32_t)((CS - 1) * 3 * BT * U * 2));
                    if (lane_ok) {
                        if (nkb > 0) { mbar_wait(sm.tmem_full, it & 1); tcgen05_fence_after(); }
                        if (threadIdx.x == CTRL_THREADS && bt == 0) stamp(c, s, 4);
                        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
                        for (int g = 0; g < 3; ++g) stage_gate<3>(trow, g * UC, g, me, row, grp, sm, nkb > 0, acc[g]);
                        fence_proxy_async_smem();
                        useful_bar_sync();
                        if (threadIdx.x == CTRL_THREADS) send_outbox<3>(sm, me);
                        if (nkb > 0) {
                            tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(sm.tmem_empty);
                        }
                        mbar_wait_cluster(sm.inbox_bar, it & 1);
#pragma unroll
                        for (int g = 0; g < 3; ++g) add_inbox<3>(sm, g, row, grp, acc[g]);
                    }
                    ++it;
                }
                float rr[8], zz[8], nn[8], gn[8];
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        rr[i] = fast_sigmoid(gi[0][i] + acc[0][i]);
                        zz[i] = fast_sigmoid(gi[1][i] + acc[1][i]);
                        gn[i] = acc[2][i] + bn[i];
                        nn[i] = fast_tanh(fmaf(rr[i], gn[i], gi[2][i]));
                        k_h[i] = fmaf(zz[i], k_h[i] - nn[i], nn[i]);        // (1-z)*n + z*h_prev
                    }
                    // the bf16 state is what the other CTAs wait for: store it first, publish, then write the rest
                    st8_bf16(p.hseq_bf + m * p.ldh + d * H + ub, k_h);
                }
                // publish step s: every epilogue thread's stores -> one release increment of the direction's counter
                if (threadIdx.x == CTRL_THREADS && bt == 0) stamp(c, s, 5);
                epi_bar_sync();                              // (the consumer fences generic->async proxy after its acquire)
                if (threadIdx.x == CTRL_THREADS) {
                    if (bt == 0) stamp(c, s, 6);
                    __threadfence();
                    atomicAdd(c.counters + d * CNT_STRIDE, 1u);
                    if (bt == 0) stamp(c, s, 7);
                }
                if (row_ok) {                                // off the critical path: nobody else reads these during the launch
                    st8(p.hseq + m * p.ldh + d * H + ub, k_h);
                    if (p.r) {
                        const size_t o = ((size_t)d * c.Tp * B + m) * H + ub;
                        st8(p.r + o, rr); st8(p.z + o, zz); st8(p.n + o, nn); st8(p.hn + o, gn);
                    }
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
};

__global__ void __launch_bounds__(THREADS, 1)
gru_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmWT, const __grid_constant__ CUtensorMap tmG, const BwdParams p) {
    constexpr int NB = UC;                           // 64 output units of the cluster
    constexpr int INBOX = (CS - 1) * BT * U * 2;
    extern __shared__ uint8_t smem_raw[];
    const Common& c = p.c;
    const int H = c.H, B = c.B;
    const int me = (int)cluster_rank();
    const int nkb_all = 3 * H / BK, nkb_max = (nkb_all + CS - 1) / CS;
    const int kb_lo = me * nkb_all / CS, nkb = (me + 1) * nkb_all / CS - kb_lo;
    const Smem sm = carve(smem_raw, nkb_max, NB, INBOX);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = blockIdx.x / c.nper;
    const int uc0 = ((blockIdx.x - d * c.nper) / CS) * UC;
    const bool rev = (d == 1) || (c.reverse0 != 0);
    const uint32_t tmem_base = setup(sm, warp, lane, nkb_max);

    if (warp == 0 && lane == 0 && nkb > 0) {
        // stationary weights: rows uc0 .. uc0+64 of this direction's W_hh^T [H, 3H], the CTA's quarter of K = 3H
        mbar_expect_tx(sm.wbar, (uint32_t)(nkb * NB * 128));
        for (int kb = 0; kb < nkb; ++kb)
            tma_load_2d(&tmWT, sm.wbar, sm.w + (size_t)kb * NB * 128, (kb_lo + kb) * BK, d * H + uc0);
    }
    if (warp < 4) {
        control_warps<NB>(sm, &tmG, warp, lane, tmem_base, c, d, kb_lo, nkb, d * 3 * H, true);
    } else {
        const int e = warp - 4, q = e & 3, grp = e >> 2;
        const bool lane_ok = q < 2;
        const int row = q * 32 + lane;
        const int ub = uc0 + me * U + grp * UPT;
        const int n_bt = (B + BT - 1) / BT;
        uint32_t it = 0;
        for (int bt = 0; bt < n_bt; ++bt) {
            const int b = bt * BT + row;
            const bool row_ok = lane_ok && b < B;
            float cr[8];                                     // dh_t * z_t carried to the next step, in registers
#pragma unroll
            for (int i = 0; i < 8; ++i) cr[i] = 0.f;
            for (int s = 0; s < c.Tp; ++s) {
                const int t = rev ? s : (c.Tp - 1 - s);              // BPTT visits time in the opposite order of the forward pass
                const int tprev = rev ? t + 1 : t - 1;               // forward-time predecessor (source of h_{t-1})
                const bool has_prev = rev ? (t + 1 < c.Tp) : (t > 0);
                const size_t m = (size_t)t * B + b;
                float dh[8], rr[8], zz[8], nn[8], gn[8], hp[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) hp[i] = 0.f;
                if (row_ok) {
                    const size_t o = ((size_t)d * c.Tp * B + m) * H + ub;
                    ld8g(p.dhseq + m * p.lddh + d * H + ub, dh);
                    ld8g(p.r + o, rr); ld8g(p.z + o, zz); ld8g(p.n + o, nn); ld8g(p.hn + o, gn);
                    if (has_prev) ld8g(p.hseq + ((size_t)tprev * B + b) * p.ldh + d * H + ub, hp);
                }
                float acc[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = 0.f;
                if (s > 0) {
                    if (threadIdx.x == CTRL_THREADS) mbar_expect_tx(sm.inbox_bar, (uint32_t)((CS - 1) * BT * U * 2));
                    if (lane_ok) {
                        if (nkb > 0) { mbar_wait(sm.tmem_full, it & 1); tcgen05_fence_after(); }
                        if (threadIdx.x == CTRL_THREADS && bt == 0) stamp(c, s, 4);
                        stage_gate<1>(tmem_base + ((uint32_t)(q * 32) << 16), 0, 0, me, row, grp, sm, nkb > 0, acc);
                        fence_proxy_async_smem();
                        useful_bar_sync();
                        if (threadIdx.x == CTRL_THREADS) send_outbox<1>(sm, me);
                        if (nkb > 0) {
                            tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(sm.tmem_empty);
                        }
                        mbar_wait_cluster(sm.inbox_bar, it & 1);
                        add_inbox<1>(sm, 0, row, grp, acc);
                    }
                    ++it;
                }
                float drt[8], dzt[8], dnt[8];
                if (row_ok) {
                    float dgn[8];
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
                    __nv_bfloat16* gh_row = p.dgh + m * p.ldg + d * 3 * H + ub;
                    st8_bf16(gh_row, drt); st8_bf16(gh_row + H, dzt); st8_bf16(gh_row + 2 * H, dgn);   // what the other CTAs wait for
                }
                if (threadIdx.x == CTRL_THREADS && bt == 0) stamp(c, s, 5);
                epi_bar_sync();                              // (the consumer fences generic->async proxy after its acquire)
                if (threadIdx.x == CTRL_THREADS) {
                    if (bt == 0) stamp(c, s, 6);
                    __threadfence();
                    atomicAdd(c.counters + d * CNT_STRIDE, 1u);
                    if (bt == 0) stamp(c, s, 7);
                }
                if (row_ok) {                                // off the critical path
                    __nv_bfloat16* gi_row = p.dgi + m * p.ldg + d * 3 * H + ub;
                    st8_bf16(gi_row, drt); st8_bf16(gi_row + H, dzt); st8_bf16(gi_row + 2 * H, dnt);
                }
            }
        }
    }
    teardown(warp, tmem_base);
}

// ---------------------------------------------------------------- host side
template <typename Kern, typename P>
static int launch_cluster_coop(Kern kern, int grid, size_t smem, const CUtensorMap& m0, const CUtensorMap& m1, const P& p, cudaStream_t s) {
    NSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = CS; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
    attrs[1].id = cudaLaunchAttributeCooperative;
    attrs[1].val.cooperative = 1;
    cfg.attrs = attrs; cfg.numAttrs = 2;
    int max_clusters = 0;
    NSD_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
    if (max_clusters * CS < grid) { set_error("gru_tc: %d CTAs in clusters of %d cannot be co-resident (max %d clusters)", grid, CS, max_clusters); return NSD_ERR_INVALID; }
    NSD_CUDA(cudaLaunchKernelEx(&cfg, kern, m0, m1, p));
    count_launch(1);
    return NSD_OK;
}

// Debug aid: NSD_GRU_TRACE=1 prints block 0's per-step event times (SM cycles relative to the step's barrier pass).
static long long* trace_begin() {
    const char* e = getenv("NSD_GRU_TRACE");
    if (!e || e[0] != '1') return nullptr;
    long long* d = nullptr;
    if (cudaMalloc(&d, sizeof(long long) * (TRACE_STEPS + 160) * 8) != cudaSuccess) return nullptr;
    cudaMemset(d, 0, sizeof(long long) * (TRACE_STEPS + 160) * 8);
    return d;
}
static void trace_end(const char* who, long long* d, cudaStream_t s, int grid = 0) {
    if (!d) return;
    static long long h[(TRACE_STEPS + 160) * 8];
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
    // step 8, all blocks, ns on the global timer, relative to the earliest barrier pass
    long long t0 = -1;
    for (int b = 0; b < grid && b < 160; ++b) { const long long v = h[(TRACE_STEPS + b) * 8]; if (v > 0 && (t0 < 0 || v < t0)) t0 = v; }
    if (t0 > 0) {
        fprintf(stderr, "  step 8 per block [ns after first barrier pass]: pass / mma_committed / epi_wake / stored / published\n");
        for (int b = 0; b < grid && b < 160; ++b) {
            const long long* r = h + (TRACE_STEPS + b) * 8;
            fprintf(stderr, "   blk %3d: %6lld %6lld %6lld %6lld %6lld\n", b, r[0] - t0, r[3] - t0, r[4] - t0, r[5] - t0, r[7] - t0);
        }
    }
}

static int check_shape(const char* who, int Tp, int B, int H, int D) {
    if (!(Tp > 0 && B > 0 && H > 0 && (D == 1 || D == 2))) { set_error("%s: bad sizes Tp=%d B=%d H=%d D=%d", who, Tp, B, H, D); return NSD_ERR_INVALID; }
    if (H % UC != 0) { set_error("%s: hidden size %d must be a multiple of %d on the tensor-core path", who, H, UC); return NSD_ERR_INVALID; }
    if (D * (H / U) > sm_count()) { set_error("%s: hidden size %d x %d directions needs %d co-resident CTAs", who, H, D, D * (H / U)); return NSD_ERR_INVALID; }
    return NSD_OK;
}
static int dbg_flag() { const char* e = getenv("NSD_GRU_DBG"); return e ? atoi(e) : 0; }

}  // namespace rtc
}  // namespace nsd

extern "C" {

size_t nsd_gru_tc_workspace(int B, int H, int D) { return 256 + sizeof(float) * (size_t)D * B * H; }

int nsd_gru_fwd_bf16(const float* gi, int ldgi, const void* w_hh_bf16, const float* b_hh, int Tp, int B, int H, int D,
                     int reverse0, float* hseq, void* hseq_bf16, int ldh, float* r, float* z, float* n, float* hn,
                     void* workspace, size_t workspace_bytes, void* stream) {
    using namespace nsd;
    using namespace nsd::rtc;
    int rc = check_shape("gru_fwd_bf16", Tp, B, H, D);
    if (rc) return rc;
    NSD_CHECK_ARG((r && z && n && hn) || (!r && !z && !n && !hn), "gru_fwd_bf16: save pointers must be all set or all NULL");
    NSD_CHECK_ARG((ldgi % 4) == 0 && (ldh % 8) == 0, "gru_fwd_bf16: ldgi must be a multiple of 4 and ldh of 8");
    if (workspace_bytes < nsd_gru_tc_workspace(B, H, D)) { set_error("gru_fwd_bf16: workspace too small"); return NSD_ERR_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    NSD_CUDA(cudaMemsetAsync(workspace, 0, 256, s));
    CUtensorMap tmW, tmH;
    rc = make_bf16_map(&tmW, w_hh_bf16, (long long)D * 3 * H, H, H, UC);
    if (rc) return rc;
    rc = make_bf16_map(&tmH, hseq_bf16, (long long)Tp * B, D * H, ldh, BT);
    if (rc) return rc;
    FwdParams p;
    long long* tr = trace_begin();
    p.c = {Tp, B, H, D, reverse0, H / U, reinterpret_cast<unsigned int*>(workspace), dbg_flag(), tr};
    p.gi = gi; p.ldgi = ldgi; p.b_hh = b_hh; p.hseq = hseq; p.hseq_bf = reinterpret_cast<__nv_bfloat16*>(hseq_bf16); p.ldh = ldh;
    p.r = r; p.z = z; p.n = n; p.hn = hn;
    const int nkb_max = (H / BK + CS - 1) / CS;
    const size_t smem = smem_bytes(nkb_max, 3 * UC, (CS - 1) * 3 * BT * U * 2);
    NSD_CHECK_ARG(smem <= 227 * 1024, "gru_fwd_bf16: hidden size %d needs %zu B of shared memory per CTA", H, smem);
    rc = launch_cluster_coop(gru_fwd_tc_kernel, D * (H / U), smem, tmW, tmH, p, s);
    trace_end("gru_fwd_bf16", tr, s, D * (H / U));
    return rc;
}

int nsd_gru_bwd_bf16(const float* dhseq, int lddh, const float* hseq, int ldh, const float* r, const float* z,
                     const float* n, const float* hn, const void* w_hhT_bf16, int Tp, int B, int H, int D, int reverse0,
                     void* dgi_bf16, void* dgh_bf16, int ldg, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace nsd;
    using namespace nsd::rtc;
    int rc = check_shape("gru_bwd_bf16", Tp, B, H, D);
    if (rc) return rc;
    NSD_CHECK_ARG((lddh % 4) == 0 && (ldh % 4) == 0 && (ldg % 8) == 0, "gru_bwd_bf16: leading dimensions must be multiples of 4 (f32) / 8 (bf16)");
    if (workspace_bytes < nsd_gru_tc_workspace(B, H, D)) { set_error("gru_bwd_bf16: workspace too small"); return NSD_ERR_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    NSD_CUDA(cudaMemsetAsync(workspace, 0, 256, s));
    CUtensorMap tmWT, tmG;
    rc = make_bf16_map(&tmWT, w_hhT_bf16, (long long)D * H, 3 * H, 3 * H, UC);
    if (rc) return rc;
    rc = make_bf16_map(&tmG, dgh_bf16, (long long)Tp * B, D * 3 * H, ldg, BT);
    if (rc) return rc;
    BwdParams p;
    long long* tr = trace_begin();
    p.c = {Tp, B, H, D, reverse0, H / U, reinterpret_cast<unsigned int*>(workspace), dbg_flag(), tr};
    p.dhseq = dhseq; p.lddh = lddh; p.hseq = hseq; p.ldh = ldh; p.r = r; p.z = z; p.n = n; p.hn = hn;
    p.dgi = reinterpret_cast<__nv_bfloat16*>(dgi_bf16); p.dgh = reinterpret_cast<__nv_bfloat16*>(dgh_bf16); p.ldg = ldg;
    const int nkb_max = (3 * H / BK + CS - 1) / CS;
    const size_t smem = smem_bytes(nkb_max, UC, (CS - 1) * BT * U * 2);
    NSD_CHECK_ARG(smem <= 227 * 1024, "gru_bwd_bf16: hidden size %d needs %zu B of shared memory per CTA", H, smem);
    rc = launch_cluster_coop(gru_bwd_tc_kernel, D * (H / U), smem, tmWT, tmG, p, s);
    trace_end("gru_bwd_bf16", tr, s, D * (H / U));
    return rc;
}

}  // extern "C"
