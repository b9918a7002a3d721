// Developer microbenchmark: the 9x9 diagonal-block step of the pivot chain (front4.cuh), variants timed with clock64 on
// one warp of an otherwise idle SM.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I islam_b200/csrc -o tools/chain_bench tools/chain_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "front4.cuh"
using namespace islam;

// V1: the shuffle-based column steps of front4.cuh (rows in lanes)
__global__ void v1(const double* A, double* out, long long* clk, int reps) {
    __shared__ double P[64 * 9];
    const int lane = threadIdx.x;
    for (int i = lane; i < 64 * 9; i += 32) P[i] = A[i];
    __syncwarp();
    long long t0 = clock64();
    double acc = 0;
    for (int r = 0; r < reps; ++r) {
        double a[9], b[9], e[9];
        for (int q = 0; q < 9; ++q) { a[q] = P[lane + 64 * q]; b[q] = P[32 + lane + 64 * q]; e[q] = (q == lane) ? 1.0 : 0.0; }
        bool ok = true;
        F4Col<0>::run(a, b, e, true, lane, ok);
        for (int q = 0; q < 9; ++q) acc += a[q] + b[q] + e[q];
        __syncwarp();
    }
    long long t1 = clock64();
    out[lane] = acc;
    if (lane == 0) clk[0] = (t1 - t0) / reps;
}

// shuffle throughput: 16 independent 64-bit shuffles per iteration
__global__ void shfl_tp(double* out, long long* clk, int reps) {
    const int lane = threadIdx.x;
    double x[16];
    for (int q = 0; q < 16; ++q) x[q] = lane + q;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
#pragma unroll
        for (int q = 0; q < 16; ++q) x[q] = __shfl_sync(0xffffffffu, x[q], (lane + q) & 31);
    long long t1 = clock64();
    double s = 0; for (int q = 0; q < 16; ++q) s += x[q];
    out[lane] = s;
    if (lane == 0) clk[1] = (t1 - t0) / reps;
}

// V2: every lane redundantly factors the 9x9 block held in registers (no shuffles): right-looking on [A; I], finished
// columns of L and rows of L^-T go to shared memory as soon as they are final; then every lane row-solves its own two rows
// with the inverse read back by broadcast loads.
template <int C> struct Col2 {
    static __device__ __forceinline__ void run(double (&A)[9][9], double (&E)[9][9], double* sL, double* sI, bool& ok) {
        const double d = A[C][C];
        if (!(d > 0.0) || !(d < 1e300)) ok = false;
        double r0;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
        const double e1 = fma(-d, r0, 1.0);
        const double p = fma(e1, e1, e1);
        const double is = f4_rsqrt(d);
        // multipliers for the rows below the pivot (A part) and the identity rows i <= C (E part)
#pragma unroll
        for (int i = C + 1; i < 9; ++i) {
            const double pa = A[i][C] * r0;
            const double m = fma(pa, p, pa);
#pragma unroll
            for (int c2 = C + 1; c2 <= i; ++c2) A[i][c2] = fma(-m, A[c2][C], A[i][c2]);
        }
#pragma unroll
        for (int i = 0; i <= C; ++i) {
            const double pe = E[i][C] * r0;
            const double m = fma(pe, p, pe);
#pragma unroll
            for (int c2 = C + 1; c2 < 9; ++c2) E[i][c2] = fma(-m, A[c2][C], E[i][c2]);
        }
#pragma unroll
        for (int i = C; i < 9; ++i) { A[i][C] *= is; }
#pragma unroll
        for (int i = 0; i <= C; ++i) { E[i][C] *= is; }
        Col2<C + 1>::run(A, E, sL, sI, ok);
    }
};
template <> struct Col2<9> { static __device__ __forceinline__ void run(double (&)[9][9], double (&)[9][9], double*, double*, bool&) {} };

__global__ void v2(const double* Ain, double* out, long long* clk, int reps) {
    __shared__ double P[64 * 9];
    __shared__ double sI[81];
    const int lane = threadIdx.x;
    for (int i = lane; i < 64 * 9; i += 32) P[i] = Ain[i];
    __syncwarp();
    long long t0 = clock64();
    double acc = 0;
    for (int r = 0; r < reps; ++r) {
        double A[9][9], E[9][9];
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int j = 0; j < 9; ++j) { A[i][j] = j <= i ? P[i + 64 * j] : 0.0; E[i][j] = i == j ? 1.0 : 0.0; }
        bool ok = true;
        Col2<0>::run(A, E, nullptr, nullptr, ok);
        // E[i][q] = (L^-T)[i][q] = Linv[q][i]; lane l < 9 publishes row l... here: every lane has everything; lane 0 stores
        if (lane < 9) {
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                double v = 0.0;
#pragma unroll
                for (int i = 0; i < 9; ++i) v = (i == lane) ? E[i][q] : v;
                sI[9 * q + lane] = v;
            }
        }
        __syncwarp();
        // row solve of this lane's two rows with the register-resident inverse: x[q] = sum_{k<=q} a[k] Linv[q][k] = sum_k a[k] E[k][q]
        double a[9], b[9], xa[9], xb[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) { a[q] = P[lane + 64 * q]; b[q] = P[32 + lane + 64 * q]; }
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int k = 0; k <= q; ++k) { s0 = fma(a[k], E[k][q], s0); s1 = fma(b[k], E[k][q], s1); }
            xa[q] = s0; xb[q] = s1;
        }
        for (int q = 0; q < 9; ++q) acc += xa[q] + xb[q] + A[q][q];
        __syncwarp();
    }
    long long t1 = clock64();
    out[lane] = acc;
    if (lane == 0) clk[2] = (t1 - t0) / reps;
}

int main() {
    double hA[64 * 9];
    // SPD-ish 9x9 block at rows 0..8 (column-major, ld 64), arbitrary rows below
    for (int j = 0; j < 9; ++j) for (int i = 0; i < 64; ++i) hA[i + 64 * j] = (i == j) ? 10.0 + i : 0.3 / (1 + abs(i - j)) * ((i * 7 + j * 3) % 5 - 2);
    for (int j = 0; j < 9; ++j) for (int i = 0; i < j; ++i) hA[i + 64 * j] = hA[j + 64 * i];
    double *dA, *dout; long long* dclk;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dout, 8 * 64); cudaMalloc(&dclk, 8 * 8);
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
    for (int it = 0; it < 2; ++it) {
        v1<<<1, 32>>>(dA, dout, dclk, 64);
        shfl_tp<<<1, 32>>>(dout, dclk, 256);
        v2<<<1, 32>>>(dA, dout, dclk, 64);
    }
    long long h[8]; double ho[32];
    cudaMemcpy(h, dclk, 64, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("err=%d  V1 shuffle block step: %lld clk   16 independent 64-bit shuffles: %lld clk   V2 redundant register block + row solve: %lld clk\n",
           (int)e, h[0], h[1], h[2]);
    return 0;
}
