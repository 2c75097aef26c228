// microbenchmark: legacy mma.sync.m16n8k16 bf16 throughput per SM on sm_100a
#include <cstdio>
#include <cuda_bf16.h>
__global__ void k(float* out, int iters, long long* cyc) {
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    for (int warps : {4, 8, 16}) {
        int iters = 2000;
        k<<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
        k<<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double macs = (double)warps * iters * 8 * 16 * 8 * 16;
        printf("warps/SM %2d: %lld cycles, %.1f MAC/cycle/SM\n", warps, h, macs / h);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
