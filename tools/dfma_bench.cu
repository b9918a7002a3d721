// Developer microbenchmark: per-SM DFMA / FFMA throughput and smem-operand DFMA rate on the box's GPU.
#include <cstdio>
#include <cuda_runtime.h>
template <typename T> __global__ void fma_chain(T* out, int iters) {
    T a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    T b = (T)1.000001, c = (T)0.5;
    for (int i = 0; i < iters; ++i) {
        a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c;
        a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
template <typename T> void run(const char* name, int blocks, int threads) {
    T* d; cudaMalloc(&d, sizeof(T) * blocks * threads);
    int iters = 20000;
    fma_chain<T><<<blocks, threads>>>(d, 100);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    fma_chain<T><<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)blocks * threads * iters * 8;
    printf("%s blocks=%d threads=%d: %.3f ms, %.1f GFMA/s total, %.2f FMA/clk/SM @1.9GHz (if blocks>=SMs: per SM = total/148)\n",
           name, blocks, threads, ms, fma / ms / 1e6, fma / (ms * 1e-3) / 1.9e9 / (blocks < 148 ? blocks : 148));
    cudaFree(d);
}
int main() {
    run<double>("f64", 1, 256); run<double>("f64", 1, 1024); run<double>("f64", 148, 1024); run<double>("f64", 592, 512);
    run<float>("f32", 1, 256); run<float>("f32", 1, 1024); run<float>("f32", 592, 512);
    return 0;
}
