// Developer microbenchmark: dependent-issue latencies on the box's GPU (clock64 around long dependent chains, one warp).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/lat_bench tools/lat_bench.cu && ./tools/lat_bench
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, double* sink, double x0, int n) {
    __shared__ double sm[64];
    double x = x0 + threadIdx.x, y = 1.0000001;
    sm[threadIdx.x & 63] = x;
    __syncwarp();
    long long t0, t1;
    // 1. dependent DFMA
    t0 = clock64();
    for (int i = 0; i < n; ++i) x = fma(x, y, 1e-9);
    t1 = clock64(); if (threadIdx.x == 0) out[0] = t1 - t0;
    // 2. dependent DMUL
    t0 = clock64();
    for (int i = 0; i < n; ++i) x = x * y;
    t1 = clock64(); if (threadIdx.x == 0) out[1] = t1 - t0;
    // 3. dependent 64-bit shuffle
    t0 = clock64();
    for (int i = 0; i < n; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    t1 = clock64(); if (threadIdx.x == 0) out[2] = t1 - t0;
    // 4. MUFU.RCP64H seed + (dependent) fma to keep it a chain
    t0 = clock64();
    for (int i = 0; i < n; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = fma(r, 1e-9, 1.5); }
    t1 = clock64(); if (threadIdx.x == 0) out[3] = t1 - t0;
    // 5. dependent LDS (pointer chase through shared memory)
    int idx = threadIdx.x & 63;
    __shared__ int nxt[64];
    nxt[threadIdx.x & 63] = (threadIdx.x + 7) & 63;
    __syncwarp();
    t0 = clock64();
    for (int i = 0; i < n; ++i) idx = nxt[idx];
    t1 = clock64(); if (threadIdx.x == 0) out[4] = t1 - t0;
    // 6. DFMA throughput, one warp, 12 independent chains
    double a[12];
    for (int q = 0; q < 12; ++q) a[q] = x + q;
    t0 = clock64();
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int q = 0; q < 12; ++q) a[q] = fma(a[q], y, 1e-9);
    t1 = clock64(); if (threadIdx.x == 0) out[5] = t1 - t0;
    for (int q = 0; q < 12; ++q) x += a[q];
    // 7. dependent DADD
    t0 = clock64();
    for (int i = 0; i < n; ++i) x = x + y;
    t1 = clock64(); if (threadIdx.x == 0) out[6] = t1 - t0;
    // 8. rsqrt seed
    t0 = clock64();
    for (int i = 0; i < n; ++i) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = fma(r, 1e-9, 1.5); }
    t1 = clock64(); if (threadIdx.x == 0) out[7] = t1 - t0;
    // 9. dependent FFMA (fp32) for comparison
    float f = (float)x, g = 1.0000001f;
    t0 = clock64();
    for (int i = 0; i < n; ++i) f = fmaf(f, g, 1e-9f);
    t1 = clock64(); if (threadIdx.x == 0) out[8] = t1 - t0;
    sink[threadIdx.x] = x + idx + f;
}
int main() {
    long long* d; double* s; cudaMalloc(&d, 80); cudaMalloc(&s, 8 * 1024);
    const int n = 4096;
    for (int threads : {32, 128, 512}) {
        k<<<1, threads>>>(d, s, 1.0, n);
        k<<<1, threads>>>(d, s, 1.0, n);
        long long h[10]; cudaMemcpy(h, d, 72, cudaMemcpyDeviceToHost);
        printf("threads=%d per-iteration clocks: DFMA %.1f  DMUL %.1f  SHFL64 %.1f  RCP64H+DFMA %.1f  LDS %.1f  12xDFMA(indep) %.1f  DADD %.1f  RSQ64H+DFMA %.1f  FFMA %.1f\n",
               threads, h[0] / (double)n, h[1] / (double)n, h[2] / (double)n, h[3] / (double)n, h[4] / (double)n, h[5] / (double)n, h[6] / (double)n, h[7] / (double)n, h[8] / (double)n);
    }
    return 0;
}
