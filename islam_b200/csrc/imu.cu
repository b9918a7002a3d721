// IMU pre-integration: /root/reference/imu_integrator.py:69-164 (IMUModule.integrate's per-frame Python loop)
// fused with pp.module.IMUPreintegrator.forward into three launches for a whole trajectory:
//   1. per-frame rotation increment  Q_f = prod_k Exp(w_k dt_k)
//   2. exclusive quaternion prefix product over frames  ->  attitude R0_f at every frame start
//   3. per-frame  dv, dp  (gravity seen through the END-of-step attitude, A.5) rotated to the world, and for
//      world mode a segmented prefix sum of velocity / position (a frame without IMU samples zeroes the
//      velocity, imu_integrator.py:134-140).
// Both modes of integrate() are covered: motion_mode=0 world-frame chain, motion_mode=1 relative deltas.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/islam_pvgo.h"
#include "lie.cuh"

using namespace islam;

namespace {

constexpr int SCAN_THREADS = 1024;
constexpr int WSCAN_THREADS = 512;

__global__ void __launch_bounds__(128)
k_imu_frame_rot(const float* __restrict__ gyro, const float* __restrict__ dt, const int* __restrict__ off, int K,
                double* __restrict__ Q) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= K) return;
    double q[4] = {0., 0., 0., 1.};
    for (int k = off[f]; k < off[f + 1]; ++k) {
        double h = dt[k];
        double w[3] = {gyro[3 * (size_t)k] * h, gyro[3 * (size_t)k + 1] * h, gyro[3 * (size_t)k + 2] * h};
        double e[4];
        so3_exp(w, e);
        q_mul(q, e, q);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) Q[4 * (size_t)f + i] = q[i];
}

// single-block exclusive scan of quaternion products: R0[f] = init * Q[0] * ... * Q[f-1]
__global__ void __launch_bounds__(SCAN_THREADS)
k_imu_scan_rot(const double* __restrict__ Q, const float* __restrict__ init, int K, double* __restrict__ R0) {
    __shared__ double sq[SCAN_THREADS][4];
    int t = threadIdx.x;
    int chunk = (K + SCAN_THREADS - 1) / SCAN_THREADS;
    int b = min(K, t * chunk), e = min(K, b + chunk);
    double q[4] = {0., 0., 0., 1.};
    for (int f = b; f < e; ++f) q_mul(q, Q + 4 * (size_t)f, q);
#pragma unroll
    for (int i = 0; i < 4; ++i) sq[t][i] = q[i];
    __syncthreads();
    // inclusive Hillis-Steele scan over thread totals (ordered product: earlier * later)
    for (int d = 1; d < SCAN_THREADS; d <<= 1) {
        double a[4], r[4];
        bool act = t >= d;
        if (act) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sq[t - d][i]; r[i] = sq[t][i]; }
            q_mul(a, r, r);
        }
        __syncthreads();
        if (act) {
#pragma unroll
            for (int i = 0; i < 4; ++i) sq[t][i] = r[i];
        }
        __syncthreads();
    }
    double p[4] = {init[3], init[4], init[5], init[6]};
    if (t > 0) q_mul(p, sq[t - 1], p);
    for (int f = b; f < e; ++f) {
#pragma unroll
        for (int i = 0; i < 4; ++i) R0[4 * (size_t)f + i] = p[i];
        q_mul(p, Q + 4 * (size_t)f, p);
    }
}

// per-frame integration with the frame-start attitude known
__global__ void __launch_bounds__(128)
k_imu_frame_integrate(const float* __restrict__ acc, const float* __restrict__ gyro, const float* __restrict__ dt,
                      const int* __restrict__ off, int K, const double* __restrict__ R0, const double* __restrict__ Q,
                      float gravity, int motion_mode, double* __restrict__ dV, double* __restrict__ dP,
                      double* __restrict__ T, float* __restrict__ pos, float* __restrict__ rot, float* __restrict__ vel) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= K) return;
    double r0[4], r0i[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) r0[i] = R0[4 * (size_t)f + i];
    q_inv(r0, r0i);
    double q[4] = {0., 0., 0., 1.};
    double dv[3] = {0., 0., 0.}, dp[3] = {0., 0., 0.}, tt = 0.;
    double gw[3] = {0., 0., (double)gravity};
    double g0[3];
    q_rot(r0i, gw, g0);                       // R0^-1 g
    int k0 = off[f], k1 = off[f + 1];
    for (int k = k0; k < k1; ++k) {
        double h = dt[k];
        double w[3] = {gyro[3 * (size_t)k] * h, gyro[3 * (size_t)k + 1] * h, gyro[3 * (size_t)k + 2] * h};
        double e[4], qn[4], qni[4], gb[3], a[3], ra[3];
        so3_exp(w, e);
        q_mul(q, e, qn);                      // dR_{k+1}
        q_inv(qn, qni);
        q_rot(qni, g0, gb);                   // (R0 dR_{k+1})^-1 g
#pragma unroll
        for (int i = 0; i < 3; ++i) a[i] = (double)acc[3 * (size_t)k + i] - gb[i];
        q_rot(q, a, ra);                      // dR_k a_k
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            dp[i] += dv[i] * h + 0.5 * ra[i] * h * h;
            dv[i] += ra[i] * h;
        }
        tt += h;
#pragma unroll
        for (int i = 0; i < 4; ++i) q[i] = qn[i];
    }
    double wv[3], wp[3];
    q_rot(r0, dv, wv);
    q_rot(r0, dp, wp);
    bool gap = (k1 == k0);
    if (motion_mode) {
        // rot = last_rot^-1 (R0 dR) = dR ; vel = R0 dv ; pos = R0 dp   (pos/vel of last_state stay zero)
#pragma unroll
        for (int i = 0; i < 3; ++i) { pos[3 * (size_t)f + i] = gap ? 0.f : (float)wp[i]; vel[3 * (size_t)f + i] = gap ? 0.f : (float)wv[i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i) rot[4 * (size_t)f + i] = (float)q[i];
    } else {
        double rq[4];
        q_mul(r0, Q + 4 * (size_t)f, rq);     // same product order as the prefix scan => consistent chain
#pragma unroll
        for (int i = 0; i < 4; ++i) rot[4 * (size_t)f + i] = (float)rq[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) { dV[3 * (size_t)f + i] = wv[i]; dP[3 * (size_t)f + i] = wp[i]; }
        T[f] = gap ? -1.0 : tt;               // negative marks "no IMU between these frames"
    }
}

// world mode: v_{f+1} = v_f + dV_f (reset to 0 on gaps), p_{f+1} = p_f + dP_f + v_f T_f ; double accumulators
__global__ void __launch_bounds__(WSCAN_THREADS)
k_imu_scan_world(const double* __restrict__ dV, const double* __restrict__ dP, const double* __restrict__ T,
                 const float* __restrict__ init, int K, float* __restrict__ pos, float* __restrict__ vel) {
    // affine state update per frame on (v, p):  v' = m v + a ;  p' = p + T v + b   with m in {0,1}
    __shared__ double sm[WSCAN_THREADS], sa[WSCAN_THREADS][3], sT[WSCAN_THREADS], sb[WSCAN_THREADS][3];
    int t = threadIdx.x;
    int chunk = (K + WSCAN_THREADS - 1) / WSCAN_THREADS;
    int b0 = min(K, t * chunk), e0 = min(K, b0 + chunk);
    double m = 1.0, a[3] = {0, 0, 0}, Ts = 0.0, b[3] = {0, 0, 0};
    for (int f = b0; f < e0; ++f) {
        bool gap = T[f] < 0.0;
        double tf = gap ? 0.0 : (double)T[f];
        // compose (existing) then frame f:  p' = p + Ts v + b ; then p'' = p' + tf v' + dP, v'' = mf v' + dV
        for (int i = 0; i < 3; ++i) b[i] += tf * a[i] + (gap ? 0.0 : (double)dP[3 * (size_t)f + i]);
        Ts += tf * m;
        if (gap) { m = 0.0; a[0] = a[1] = a[2] = 0.0; }
        else for (int i = 0; i < 3; ++i) a[i] += (double)dV[3 * (size_t)f + i];
    }
    sm[t] = m; sT[t] = Ts;
    for (int i = 0; i < 3; ++i) { sa[t][i] = a[i]; sb[t][i] = b[i]; }
    __syncthreads();
    for (int d = 1; d < WSCAN_THREADS; d <<= 1) {
        double m2 = 0, T2 = 0, a2[3], b2[3];
        bool act = t >= d;
        if (act) {
            // earlier = (m1,a1,T1,b1) at t-d, later = (m,a,T,b) at t:  v'' = m(m1 v + a1) + a ; p'' = p + T1 v + b1 + T (m1 v + a1) + b
            double m1 = sm[t - d], T1 = sT[t - d];
            m2 = sm[t] * m1;
            T2 = T1 + sT[t] * m1;
            for (int i = 0; i < 3; ++i) {
                a2[i] = sm[t] * sa[t - d][i] + sa[t][i];
                b2[i] = sb[t - d][i] + sT[t] * sa[t - d][i] + sb[t][i];
            }
        }
        __syncthreads();
        if (act) {
            sm[t] = m2; sT[t] = T2;
            for (int i = 0; i < 3; ++i) { sa[t][i] = a2[i]; sb[t][i] = b2[i]; }
        }
        __syncthreads();
    }
    double v[3], p[3];
    for (int i = 0; i < 3; ++i) { v[i] = (double)init[7 + i]; p[i] = (double)init[i]; }
    if (t > 0) {
        double m1 = sm[t - 1], T1 = sT[t - 1];
        for (int i = 0; i < 3; ++i) {
            double vi = v[i];
            v[i] = m1 * vi + sa[t - 1][i];
            p[i] = p[i] + T1 * vi + sb[t - 1][i];
        }
    }
    for (int f = b0; f < e0; ++f) {
        bool gap = T[f] < 0.0;
        double tf = gap ? 0.0 : (double)T[f];
        for (int i = 0; i < 3; ++i) {
            if (!gap) p[i] += tf * v[i] + (double)dP[3 * (size_t)f + i];
            v[i] = gap ? 0.0 : v[i] + (double)dV[3 * (size_t)f + i];
            pos[3 * (size_t)f + i] = (float)p[i];
            vel[3 * (size_t)f + i] = (float)v[i];
        }
    }
}

}  // namespace

extern "C" int64_t islam_imu_workspace_bytes(int32_t S, int32_t K) {
    (void)S;
    return (int64_t)sizeof(double) * 16 * (int64_t)(K > 0 ? K : 0) + 256;
}

extern "C" int islam_imu_preintegrate(const float* acc, const float* gyro, const float* dt, int32_t S,
                                      const int32_t* offsets, int32_t K, const float* init, float gravity,
                                      int32_t motion_mode, float* pos, float* rot, float* vel, void* workspace,
                                      void* stream) {
    if (K <= 0) return 0;
    if (!acc || !gyro || !dt || !offsets || !init || !pos || !rot || !vel || !workspace || S < 0) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    double* ws = (double*)workspace;     // float32 samples in, float64 accumulation (a 45 000-sample quaternion chain
    double* Q = ws;                      // K x 4     in float32 drifts by ~1e-5 rad, which gravity turns into metres)
    double* R0 = Q + 4 * (size_t)K;      // K x 4
    double* dV = R0 + 4 * (size_t)K;     // K x 3
    double* dP = dV + 3 * (size_t)K;     // K x 3
    double* T = dP + 3 * (size_t)K;      // K
    int nb = (K + 127) / 128;
    k_imu_frame_rot<<<nb, 128, 0, s>>>(gyro, dt, offsets, K, Q);
    k_imu_scan_rot<<<1, SCAN_THREADS, 0, s>>>(Q, init, K, R0);
    k_imu_frame_integrate<<<nb, 128, 0, s>>>(acc, gyro, dt, offsets, K, R0, Q, gravity, motion_mode, dV, dP, T, pos, rot, vel);
    if (!motion_mode) k_imu_scan_world<<<1, WSCAN_THREADS, 0, s>>>(dV, dP, T, init, K, pos, vel);
    return (int)cudaGetLastError();
}
