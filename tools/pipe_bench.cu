// Developer microbenchmark: how much does a dependent DFMA chain (the pivot chain of a front) slow down when the other warps of
// the SM keep the fp64 pipe busy with (a) independent DFMAs, (b) FP64 tensor-core DMMAs (mma.sync.m8n8k4.f64)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/pipe_bench tools/pipe_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, double* sink, int mode, int nload, int n) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ volatile int stop;
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    if (warp == 0) {
        double x = 1.0 + lane, y = 1.0000001;
        long long t0 = clock64();
        for (int i = 0; i < n; ++i) x = fma(x, y, 1e-9);
        long long t1 = clock64();
        if (lane == 0) { out[0] = t1 - t0; stop = 1; }
        sink[threadIdx.x] = x;
        return;
    }
    if (warp > nload) return;
    double acc[16];
    for (int q = 0; q < 16; ++q) acc[q] = lane + q;
    double a = 1.0000001 + lane * 1e-9, b = 0.5;
    long long cnt = 0;
    while (!stop) {
        if (mode == 0) {
#pragma unroll
            for (int q = 0; q < 16; ++q) acc[q] = fma(acc[q], a, b);
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(acc[2 * q]), "+d"(acc[2 * q + 1]) : "d"(a), "d"(b));
        }
        ++cnt;
    }
    double s = 0; for (int q = 0; q < 16; ++q) s += acc[q];
    sink[threadIdx.x] = s;
    if (lane == 0) out[warp] = cnt;
}
int main() {
    long long* d; double* s; cudaMalloc(&d, 8 * 64); cudaMalloc(&s, 8 * 1024);
    const int n = 20000;
    for (int mode = 0; mode < 2; ++mode)
        for (int nload : {0, 1, 3, 6, 7, 15}) {
            cudaMemset(d, 0, 8 * 64);
            k<<<1, 32 * 16>>>(d, s, mode, nload, n);
            long long h[16]; cudaMemcpy(h, d, 8 * 16, cudaMemcpyDeviceToHost);
            long long iters = 0; for (int w = 1; w <= nload; ++w) iters += h[w];
            double fma_per_clk = mode == 0 ? iters * 16.0 * 32 / h[0] : iters * 8.0 * 256 / h[0];
            printf("%s load warps %2d: chain %.2f clk per dependent DFMA; load throughput %.1f FMA/clk\n", mode == 0 ? "DFMA" : "DMMA", nload, h[0] / (double)n, fma_per_clk);
        }
    printf("err %d\n", (int)cudaDeviceSynchronize());
    return 0;
}
