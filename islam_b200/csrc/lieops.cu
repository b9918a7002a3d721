// Elementwise LieTensor maps with PyPose's left-tangent autograd convention: the gradient a
// backward returns for a group-valued input is d L / d(delta) for X <- Exp(delta) X, stored in the leading
// 6 (SE3) or 3 (SO3) slots of an embedding-sized row (7 / 4), last slot zero.  These back the LieTensor shim
// (islam_b200/pypose_compat) on CUDA tensors: Exp/Log/Inv/@ used at /root/reference/pvgo.py:38-39,47-48,72-73,
// Datasets/transformation.py:72-124 and train.py:215.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/islam_pvgo.h"
#include "lie.cuh"

using namespace islam;

namespace {

constexpr int T = 128;
#define IDX int64_t i = (int64_t)blockIdx.x * T + threadIdx.x; if (i >= n) return

// 6x6 se3 left Jacobian pieces: Jl(phi) and Q(xi);  Jl6 = [[Jl, Q],[0, Jl]]
__device__ __forceinline__ void ld(const float* p, float* x, int k) { for (int q = 0; q < k; ++q) x[q] = p[q]; }
__device__ __forceinline__ void stv(float* p, const float* x, int k) { for (int q = 0; q < k; ++q) p[q] = x[q]; }
__device__ __forceinline__ void matT_vec(const float* M, const float* v, float* o) {   // o = M^T v
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = M[c] * v[0] + M[3 + c] * v[1] + M[6 + c] * v[2];
}
__device__ __forceinline__ void mat_vec(const float* M, const float* v, float* o) {
#pragma unroll
    for (int r = 0; r < 3; ++r) o[r] = M[3 * r] * v[0] + M[3 * r + 1] * v[1] + M[3 * r + 2] * v[2];
}
// Ad(X)^T g for SE3: Ad = [[R, [t]x R],[0, R]]  =>  Ad^T g = [R^T g_t ; -R^T [t]x g_t + R^T g_r] ... ([t]x R)^T = -R^T [t]x
__device__ __forceinline__ void adjT_se3(const float* X, const float* g, float* o) {
    float qi[4], a[3], b[3], c[3];
    q_inv(X + 3, qi);
    q_rot(qi, g, a);                 // R^T g_t
    cross3(X, g, c);                 // t x g_t
    q_rot(qi, c, b);                 // R^T (t x g_t)
    float d[3];
    q_rot(qi, g + 3, d);             // R^T g_r
    o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
    o[3] = d[0] - b[0]; o[4] = d[1] - b[1]; o[5] = d[2] - b[2];
}

__global__ void k_exp(int group, const float* x, float* y, int64_t n) {
    IDX;
    if (group == ISLAM_SE3) { float a[6], o[7]; ld(x + 6 * i, a, 6); se3_exp(a, o); stv(y + 7 * i, o, 7); }
    else { float a[3], o[4]; ld(x + 3 * i, a, 3); so3_exp(a, o); stv(y + 4 * i, o, 4); }
}
__global__ void k_log(int group, const float* x, float* y, int64_t n) {
    IDX;
    if (group == ISLAM_SE3) { float a[7], o[6], Ji[9]; ld(x + 7 * i, a, 7); se3_log(a, o, Ji); stv(y + 6 * i, o, 6); }
    else { float a[4], o[3]; ld(x + 4 * i, a, 4); so3_log(a, o); stv(y + 3 * i, o, 3); }
}
__global__ void k_inv(int group, const float* x, float* y, int64_t n) {
    IDX;
    if (group == ISLAM_SE3) { float a[7], o[7]; ld(x + 7 * i, a, 7); se3_inv(a, o); stv(y + 7 * i, o, 7); }
    else { float a[4], o[4]; ld(x + 4 * i, a, 4); q_inv(a, o); stv(y + 4 * i, o, 4); }
}
__global__ void k_mul(int group, const float* a_, const float* b_, float* y, int64_t n) {
    IDX;
    if (group == ISLAM_SE3) { float a[7], b[7], o[7]; ld(a_ + 7 * i, a, 7); ld(b_ + 7 * i, b, 7); se3_mul(a, b, o); stv(y + 7 * i, o, 7); }
    else { float a[4], b[4], o[4]; ld(a_ + 4 * i, a, 4); ld(b_ + 4 * i, b, 4); q_mul(a, b, o); stv(y + 4 * i, o, 4); }
}
__global__ void k_act(int group, const float* x, const float* p_, float* y, int64_t n) {
    IDX;
    float p[3], o[3];
    ld(p_ + 3 * i, p, 3);
    if (group == ISLAM_SE3) { float a[7]; ld(x + 7 * i, a, 7); q_rot(a + 3, p, o); o[0] += a[0]; o[1] += a[1]; o[2] += a[2]; }
    else { float a[4]; ld(x + 4 * i, a, 4); q_rot(a, p, o); }
    stv(y + 3 * i, o, 3);
}

// Exp backward: Exp(x + dx) = Exp(Jl(x) dx) Exp(x)  =>  gx = Jl(x)^T gy[:dim]
__global__ void k_exp_bwd(int group, const float* x, const float* gy, float* gx, int64_t n) {
    IDX;
    if (group == ISLAM_SE3) {
        float a[6], g[6], Jl[9], Q[9], o[6], t1[3], t2[3];
        ld(x + 6 * i, a, 6); ld(gy + 7 * i, g, 6);
        so3_Jl(a + 3, Jl);
        se3_Q(a, Q);
        matT_vec(Jl, g, o);                 // d tau: Jl^T g_t
        matT_vec(Q, g, t1);                 // d phi: Q^T g_t + Jl^T g_r
        matT_vec(Jl, g + 3, t2);
        o[3] = t1[0] + t2[0]; o[4] = t1[1] + t2[1]; o[5] = t1[2] + t2[2];
        stv(gx + 6 * i, o, 6);
    } else {
        float a[3], g[3], Jl[9], o[3];
        ld(x + 3 * i, a, 3); ld(gy + 4 * i, g, 3);
        so3_Jl(a, Jl);
        matT_vec(Jl, g, o);
        stv(gx + 3 * i, o, 3);
    }
}
// Log backward: Log(Exp(d) X) = x + Jl^-1(x) d  =>  gX[:dim] = Jl^-1(x)^T gy
__global__ void k_log_bwd(int group, const float* y, const float* gy, float* gx, int64_t n) {
    IDX;
    if (group == ISLAM_SE3) {
        float a[6], g[6], Ji[9], Q[9], T1[9], B[9], o[7], t1[3], t2[3];
        ld(y + 6 * i, a, 6); ld(gy + 6 * i, g, 6);
        so3_Jl_inv(a + 3, Ji);
        se3_Q(a, Q);
        mat3_mul(Ji, Q, T1);
        mat3_mul(T1, Ji, B);                // Ji Q Ji ; upper-right block of Jl6^-1 is -B
        matT_vec(Ji, g, o);
        matT_vec(B, g, t1);
        matT_vec(Ji, g + 3, t2);
        o[3] = t2[0] - t1[0]; o[4] = t2[1] - t1[1]; o[5] = t2[2] - t1[2];
        o[6] = 0.f;
        stv(gx + 7 * i, o, 7);
    } else {
        float a[3], g[3], Ji[9], o[4];
        ld(y + 3 * i, a, 3); ld(gy + 3 * i, g, 3);
        so3_Jl_inv(a, Ji);
        matT_vec(Ji, g, o);
        o[3] = 0.f;
        stv(gx + 4 * i, o, 4);
    }
}
// Inv backward: (Exp(d) X)^-1 = Exp(-Ad(X^-1) d) X^-1  =>  gX = -Ad(Y)^T gY,  Y = X^-1
__global__ void k_inv_bwd(int group, const float* y, const float* gy, float* gx, int64_t n) {
    IDX;
    if (group == ISLAM_SE3) {
        float Y[7], g[6], o[7];
        ld(y + 7 * i, Y, 7); ld(gy + 7 * i, g, 6);
        adjT_se3(Y, g, o);
#pragma unroll
        for (int k = 0; k < 6; ++k) o[k] = -o[k];
        o[6] = 0.f;
        stv(gx + 7 * i, o, 7);
    } else {
        float Y[4], Yi[4], g[3], o[4];
        ld(y + 4 * i, Y, 4); ld(gy + 4 * i, g, 3);
        q_inv(Y, Yi);
        q_rot(Yi, g, o);                    // R(Y)^T g
        o[0] = -o[0]; o[1] = -o[1]; o[2] = -o[2]; o[3] = 0.f;
        stv(gx + 4 * i, o, 4);
    }
}
// Mul backward: Exp(d) A B -> gA = gY ;  A Exp(d) B = Exp(Ad(A) d) A B -> gB = Ad(A)^T gY
__global__ void k_mul_bwd(int group, const float* a_, const float* gy, float* ga, float* gb, int64_t n) {
    IDX;
    if (group == ISLAM_SE3) {
        float A[7], g[7], o[7];
        ld(a_ + 7 * i, A, 7); ld(gy + 7 * i, g, 6);
        g[6] = 0.f;
        if (ga) stv(ga + 7 * i, g, 7);
        if (gb) { adjT_se3(A, g, o); o[6] = 0.f; stv(gb + 7 * i, o, 7); }
    } else {
        float A[4], Ai[4], g[4], o[4];
        ld(a_ + 4 * i, A, 4); ld(gy + 4 * i, g, 3);
        g[3] = 0.f;
        if (ga) stv(ga + 4 * i, g, 4);
        if (gb) { q_inv(A, Ai); q_rot(Ai, g, o); o[3] = 0.f; stv(gb + 4 * i, o, 4); }
    }
}
// Act backward: y = X p ; d y / d delta = [I, -[y]x] (SE3) or -[y]x (SO3) ; d y / d p = R
__global__ void k_act_bwd(int group, const float* x, const float* p_, const float* gy, float* gx, float* gp, int64_t n) {
    IDX;
    float p[3], g[3], yv[3], c[3];
    ld(p_ + 3 * i, p, 3); ld(gy + 3 * i, g, 3);
    if (group == ISLAM_SE3) {
        float A[7], Ai[4], o[7];
        ld(x + 7 * i, A, 7);
        q_rot(A + 3, p, yv); yv[0] += A[0]; yv[1] += A[1]; yv[2] += A[2];
        cross3(yv, g, c);
        o[0] = g[0]; o[1] = g[1]; o[2] = g[2]; o[3] = c[0]; o[4] = c[1]; o[5] = c[2]; o[6] = 0.f;
        if (gx) stv(gx + 7 * i, o, 7);
        if (gp) { q_inv(A + 3, Ai); q_rot(Ai, g, c); stv(gp + 3 * i, c, 3); }
    } else {
        float A[4], Ai[4], o[4];
        ld(x + 4 * i, A, 4);
        q_rot(A, p, yv);
        cross3(yv, g, c);
        o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3] = 0.f;
        if (gx) stv(gx + 4 * i, o, 4);
        if (gp) { q_inv(A, Ai); q_rot(Ai, g, c); stv(gp + 3 * i, c, 3); }
    }
}

// ordered prefix product along the leading dimension (pp.cumprod): left = 1: y_i = x_i * y_{i-1}; left = 0: y_i = y_{i-1} * x_i.
// One block; float64 accumulation (a 5 000-element float32 chain would drift by ~1e-5).  Datasets/transformation.py:100-113
// (motion2pose_pypose) is this scan with left = 0.
constexpr int CP_THREADS = 512;
template <int GROUP>
__global__ void __launch_bounds__(CP_THREADS) k_cumprod(const float* __restrict__ x, float* __restrict__ y, int64_t n, int left) {
    constexpr int W = GROUP == ISLAM_SE3 ? 7 : 4;
    __shared__ double sq[CP_THREADS][W];
    const int t = threadIdx.x;
    const int64_t chunk = (n + CP_THREADS - 1) / CP_THREADS;
    const int64_t b = min(n, (int64_t)t * chunk), e = min(n, b + chunk);
    auto mul = [&](const double* A, const double* B, double* O) {      // O = left ? B*A : A*B  (A earlier, B later)
        if (GROUP == ISLAM_SE3) { if (left) se3_mul(B, A, O); else se3_mul(A, B, O); }
        else { if (left) q_mul(B, A, O); else q_mul(A, B, O); }
    };
    double acc[W], cur[W], tmp[W];
#pragma unroll
    for (int k = 0; k < W; ++k) acc[k] = (k == W - 1) ? 1.0 : 0.0;
    for (int64_t i = b; i < e; ++i) {
#pragma unroll
        for (int k = 0; k < W; ++k) cur[k] = (double)x[W * i + k];
        mul(acc, cur, tmp);
#pragma unroll
        for (int k = 0; k < W; ++k) acc[k] = tmp[k];
    }
#pragma unroll
    for (int k = 0; k < W; ++k) sq[t][k] = acc[k];
    __syncthreads();
    for (int d = 1; d < CP_THREADS; d <<= 1) {
        double a[W], r[W];
        const bool act = t >= d;
        if (act) {
#pragma unroll
            for (int k = 0; k < W; ++k) { a[k] = sq[t - d][k]; r[k] = sq[t][k]; }
            mul(a, r, tmp);
        }
        __syncthreads();
        if (act) {
#pragma unroll
            for (int k = 0; k < W; ++k) sq[t][k] = tmp[k];
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < W; ++k) acc[k] = (t > 0) ? sq[t - 1][k] : ((k == W - 1) ? 1.0 : 0.0);
    for (int64_t i = b; i < e; ++i) {
#pragma unroll
        for (int k = 0; k < W; ++k) cur[k] = (double)x[W * i + k];
        mul(acc, cur, tmp);
#pragma unroll
        for (int k = 0; k < W; ++k) { acc[k] = tmp[k]; y[W * i + k] = (float)tmp[k]; }
    }
}

inline int chk(int group, int64_t n) { return (group != ISLAM_SE3 && group != ISLAM_SO3) || n < 0 ? -1 : 0; }
inline unsigned grid(int64_t n) { return (unsigned)((n + T - 1) / T); }

}  // namespace

#define LAUNCH(kern, ...)                                                  \
    if (chk(group, n)) return -1;                                          \
    if (n == 0) return 0;                                                  \
    kern<<<grid(n), T, 0, (cudaStream_t)stream>>>(group, __VA_ARGS__, n);  \
    return (int)cudaGetLastError()

extern "C" int islam_lie_exp(int32_t group, const float* x, float* y, int64_t n, void* stream) { LAUNCH(k_exp, x, y); }
extern "C" int islam_lie_log(int32_t group, const float* x, float* y, int64_t n, void* stream) { LAUNCH(k_log, x, y); }
extern "C" int islam_lie_inv(int32_t group, const float* x, float* y, int64_t n, void* stream) { LAUNCH(k_inv, x, y); }
extern "C" int islam_lie_mul(int32_t group, const float* a, const float* b, float* y, int64_t n, void* stream) { LAUNCH(k_mul, a, b, y); }
extern "C" int islam_lie_act(int32_t group, const float* x, const float* p, float* y, int64_t n, void* stream) { LAUNCH(k_act, x, p, y); }
extern "C" int islam_lie_exp_bwd(int32_t group, const float* x, const float* gy, float* gx, int64_t n, void* stream) { LAUNCH(k_exp_bwd, x, gy, gx); }
extern "C" int islam_lie_log_bwd(int32_t group, const float* y, const float* gy, float* gx, int64_t n, void* stream) { LAUNCH(k_log_bwd, y, gy, gx); }
extern "C" int islam_lie_inv_bwd(int32_t group, const float* y, const float* gy, float* gx, int64_t n, void* stream) { LAUNCH(k_inv_bwd, y, gy, gx); }
extern "C" int islam_lie_mul_bwd(int32_t group, const float* a, const float* gy, float* ga, float* gb, int64_t n, void* stream) { LAUNCH(k_mul_bwd, a, gy, ga, gb); }
extern "C" int islam_lie_act_bwd(int32_t group, const float* x, const float* p, const float* gy, float* gx, float* gp, int64_t n, void* stream) { LAUNCH(k_act_bwd, x, p, gy, gx, gp); }

extern "C" int islam_lie_cumprod(int32_t group, const float* x, float* y, int64_t n, int32_t left, void* stream) {
    if (chk(group, n) || !x || !y) return -1;
    if (n == 0) return 0;
    if (group == ISLAM_SE3) k_cumprod<ISLAM_SE3><<<1, CP_THREADS, 0, (cudaStream_t)stream>>>(x, y, n, left);
    else k_cumprod<ISLAM_SO3><<<1, CP_THREADS, 0, (cudaStream_t)stream>>>(x, y, n, left);
    return (int)cudaGetLastError();
}
