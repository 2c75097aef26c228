// microbenchmark: tcgen05.mma with the A operand in TMEM (weights stationary in tensor memory, loaded with tcgen05.st),
// B from shared memory (K-major, 128B swizzle).  Checks the result against a host reference and times one K=16 slab
// for several N, in TS mode (A in TMEM) and SS mode (A in shared memory).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/umma_ts_bench scratch/umma_ts_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>

constexpr int KTOT = 256;              // reduction length: 16 slabs of 16
constexpr int M = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__host__ __device__ inline float aval(int m, int k) { return (float)(((m * 7 + k * 3) % 13) - 6) * 0.125f; }
__host__ __device__ inline float bval(int n, int k) { return (float)(((n * 5 + k * 11) % 9) - 4) * 0.25f; }

// byte offset of element (row, k) in a K-major SW128 tile stack: k-block kb (64 wide) = rows x 128 B, 8-row groups of 1024 B
__device__ __forceinline__ uint32_t sw128_off(int rows, int row, int k) {
    const int kb = k >> 6, kk = k & 63;
    const int chunk = (kk >> 3) ^ (row & 7);
    return (uint32_t)(kb * rows * 128 + (row >> 3) * 1024 + (row & 7) * 128 + chunk * 16 + (kk & 7) * 2);
}

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) bench(float* out, long long* cyc, int iters) {
    extern __shared__ uint8_t raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sB = base;                               // [KTOT/64][N rows][128 B]
    uint8_t* sA = base + (KTOT / 64) * 256 * 128;     // [KTOT/64][128 rows][128 B]  (SS mode)
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < N * KTOT; i += 128) {
        const int n = i / KTOT, k = i % KTOT;
        *reinterpret_cast<__nv_bfloat16*>(sB + sw128_off(N, n, k)) = __float2bfloat16(bval(n, k));
    }
    for (int i = threadIdx.x; i < M * KTOT; i += 128) {
        const int m = i / KTOT, k = i % KTOT;
        *reinterpret_cast<__nv_bfloat16*>(sA + sw128_off(M, m, k)) = __float2bfloat16(aval(m, k));
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = slot;
    const uint32_t tA = tb + 256;                      // A: 128 lanes x KTOT/2 = 128 columns, at column 256
    // A into TMEM: thread = lane (row) m; column c holds k = 2c (low half), 2c+1 (high half)
    {
        const int m = warp * 32 + lane;
        for (int c0 = 0; c0 < KTOT / 2; c0 += 8) {
            uint32_t v[8];
            for (int i = 0; i < 8; ++i) {
                __nv_bfloat162 p = __floats2bfloat162_rn(aval(m, 2 * (c0 + i)), aval(m, 2 * (c0 + i) + 1));
                v[i] = *reinterpret_cast<uint32_t*>(&p);
            }
            const uint32_t ta = tA + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         ::"r"(ta), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = idesc_bf16(M, N);
        uint32_t phase = 0;
        for (int rep = 0; rep < 2; ++rep) {            // rep 0 warms up, rep 1 is timed
            t0 = clock64();
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int s = 0; s < KTOT / 16; ++s) {
                    const uint64_t bdesc = kmajor_desc(smem_u32(sB + (s >> 2) * N * 128)) + (uint64_t)(2 * (s & 3));
                    const uint32_t acc = (it | s) != 0;
                    if (TS) {
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                     ::"r"(tb), "r"(tA + (uint32_t)(s * 8)), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
                    } else {
                        const uint64_t adesc = kmajor_desc(smem_u32(sA + (s >> 2) * M * 128)) + (uint64_t)(2 * (s & 3));
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                     ::"r"(tb), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
            phase ^= 1;
            t1 = clock64();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // D: lane m, columns 0..N-1
    if (blockIdx.x == 0) {
        for (int c0 = 0; c0 < N; c0 += 8) {
            uint32_t r[8];
            const uint32_t ta = tb + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(ta));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 8; ++i) out[(warp * 32 + lane) * 256 + c0 + i] = __uint_as_float(r[i]);
        }
        if (threadIdx.x == 0) *cyc = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

template <int N, bool TS>
void run(float* dout, long long* dcyc, int iters) {
    const size_t smem = (KTOT / 64) * 256 * 128 + (KTOT / 64) * 128 * 128 + 2048;
    cudaFuncSetAttribute(bench<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemset(dout, 0, 128 * 256 * 4);
    bench<N, TS><<<148, 128, smem>>>(dout, dcyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d %s: %s\n", N, TS ? "TS" : "SS", cudaGetErrorString(e)); exit(1); }
    std::vector<float> h(128 * 256);
    long long c;
    cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < KTOT; ++k) ref += (double)aval(m, k) * bval(n, k);
            ref *= iters;                              // the timed rep accumulates `iters` identical products from zero
            const double err = fabs(h[m * 256 + n] - ref) / (1.0 + fabs(ref));
            if (err > maxerr) maxerr = err;
        }
    printf("M=128 N=%3d %s: %8.1f cycles per K=16 slab (%lld cycles / %d MMAs), max rel err %.3g\n", N, TS ? "TS" : "SS",
           (double)c / (iters * (KTOT / 16)), c, iters * (KTOT / 16), maxerr);
}

int main() {
    float* dout; long long* dcyc;
    cudaMalloc(&dout, 128 * 256 * 4); cudaMalloc(&dcyc, 8);
    const int iters = 64;
    run<16, true>(dout, dcyc, iters);
    run<32, true>(dout, dcyc, iters);
    run<64, true>(dout, dcyc, iters);
    run<128, true>(dout, dcyc, iters);
    run<256, true>(dout, dcyc, iters);
    run<16, false>(dout, dcyc, iters);
    run<32, false>(dout, dcyc, iters);
    run<64, false>(dout, dcyc, iters);
    run<128, false>(dout, dcyc, iters);
    run<256, false>(dout, dcyc, iters);
    // short bursts, as in one recurrence step: 16 and 32 MMAs issued back to back, then a commit
    run<32, true>(dout, dcyc, 1);
    run<32, true>(dout, dcyc, 2);
    run<64, true>(dout, dcyc, 1);
    run<64, false>(dout, dcyc, 1);
    return 0;
}
