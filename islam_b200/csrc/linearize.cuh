// Kernel family 1 — per-factor SE(3) residuals + Jacobian blocks and deterministic assembly of the
// block-sparse normal equations  H = J^T W J (float64 9x9 blocks),  g = J^T W r.
//
// Restates /root/reference/pvgo.py:26-64 (PoseVelGraph.forward) together with the Jacobian blocks that
// PyPose's autograd produces for it and the information weights of pvgo.py:125-129.
// Residual/Jacobian arithmetic is float32 (the reference's dtype); products and sums that form H and g
// are float64 because the damped system has cond ~1e8 (DESIGN.md "precision").
#pragma once
#include "common.cuh"
#include "lie.cuh"

namespace islam {

constexpr int LIN_THREADS = 128;

struct ProblemView {
    int N, E, M;
    const int* ei;        // [E] first endpoint (pvgo.py:36  nodes[edges[:,0]])
    const int* ej;        // [E] second endpoint
    const float* Z;       // [E,7] vo_motions
    const float* drot;    // [M,4]
    const float* dtrans;  // [M,3]
    const float* dvel;    // [M,3]
    const float* dt;      // [M]
    const int* edge_owner;  // [E] window that owns the factor (multi-GPU) or nullptr
    const int* pair_owner;  // [M]
    int part;
    const double* w;        // [5] information scalars w0..w3 (pvgo.py:125-129) and the reprojection one (pvgo.py:131) in device memory
    // optional sparse reprojection factor (pvgo.py:53-61, dense_ba.py:276-305); rp_n = 0: absent
    const float* rp_pts;    // [M, rp_n, 3] camera-frame points of pose i
    const float* rp_tgt;    // [M, rp_n, 2] target pixels in the camera at pose i+1
    const float* rp_cal;    // [11] fx, fy, cx, cy, rgb2imu pose (7)
    int rp_n;
};

struct LinBuffers {
    float* r_vo;     // [E,6]  pgerr
    float* J_vo;     // [E,18] Mm (3x3) then K (3x3):  J_j = [[Mm, K],[0, Mm]],  J_i = -J_j
    double* S_vo;    // [E,36] w0 * J^T J
    double* q_vo;    // [E,6]  w0 * J^T r
    float* r_imu;    // [M,9]  adjvelerr(3), imuroterr(3), transvelerr(3)
    float* J_rot;    // [M,9]  Jl^-1(r) dR^T Ri^T
    double* loss_part;   // [nblk_vo + nblk_imu + nblk_rp] partial sums of r^2 (unweighted — PyPose model.loss)
    float* r_rp;     // [M, 2 rp_n] reprojection residuals
    double* S_rp;    // [M,36] UNWEIGHTED sum over the pair's points of J^T J (J = d r / d delta_i; d r / d delta_{i+1} = -J)
    double* q_rp;    // [M,6]  unweighted J^T r
};

__device__ __forceinline__ void load7(const float* p, float* x) {
#pragma unroll
    for (int i = 0; i < 7; ++i) x[i] = p[i];
}

// r = Log(Z^-1 Xi^-1 Xj);  Mm = Jl^-1(phi) R(A),  K = (Jl^-1 [tA]x - Jl^-1 Q Jl^-1) R(A),  A = Z^-1 Xi^-1
__device__ __forceinline__ void vo_factor(const float* Xi, const float* Xj, const float* Zm, float* r, float* Mm,
                                          float* K) {
    float Zi[7], Xii[7], A[7], Eerr[7], Ji[9];
    se3_inv(Zm, Zi);
    se3_inv(Xi, Xii);
    se3_mul(Zi, Xii, A);
    se3_mul(A, Xj, Eerr);
    se3_log(Eerr, r, Ji);
    if (Mm == nullptr) return;
    float R[9], Q[9], T1[9], T2[9];
    q_matrix(A + 3, R);
    mat3_mul(Ji, R, Mm);
    se3_Q(r, Q);
    mat3_mul(Ji, Q, T1);
    mat3_mul(T1, Ji, T2);     // Ji Q Ji
    float tx[9] = {0, -A[2], A[1], A[2], 0, -A[0], -A[1], A[0], 0};
    mat3_mul(Ji, tx, T1);     // Ji [t]x
#pragma unroll
    for (int i = 0; i < 9; ++i) T1[i] -= T2[i];
    mat3_mul(T1, R, K);
}

// r = Log(dR^-1 Ri^-1 Rj);  Jrot = Jl^-1(r) R(dR^-1 Ri^-1)
__device__ __forceinline__ void rot_factor(const float* qi, const float* qj, const float* dq, float* r, float* Jr) {
    float a[4], b[4], c[4], e[4];
    q_inv(dq, a);
    q_inv(qi, b);
    q_mul(a, b, c);
    q_mul(c, qj, e);
    so3_log(e, r);
    if (Jr == nullptr) return;
    float Ji[9], R[9];
    so3_Jl_inv(r, Ji);
    q_matrix(c, R);
    mat3_mul(Ji, R, Jr);
}

// ---------------------------------------------------------------------------------------------- VO factors
// mode 0: linearise at the current state (writes r, J, S, q, loss partials)
// mode 1: trial evaluation — residuals at the trial state (loss partials) and the unweighted
//         (J D)^T (2 r + J D) term of TrustRegion.update from the stored linearisation
template <int MODE>
__device__ __forceinline__ void
vo_block(int blk, const LMState* __restrict__ st, const float* __restrict__ nodes0, const float* __restrict__ nodes1,
     ProblemView pv, LinBuffers lb, const double* __restrict__ D, double* __restrict__ part_out, int force) {
    if (!force) {
        if (!st->active) return;
        if (MODE == 0 && !st->do_lin) return;
    }
    int cur = st->cur;
    const float* nodes = (MODE == 0) ? (cur ? nodes1 : nodes0) : (cur ? nodes0 : nodes1);
    __shared__ double sh[LIN_THREADS / 32];
    __shared__ double sh2[LIN_THREADS / 32];
    int e = blk * LIN_THREADS + threadIdx.x;
    double lsum = 0.0, qsum = 0.0;
    bool mine = e < pv.E && (pv.edge_owner == nullptr || pv.edge_owner[e] == pv.part);
    if (mine) {
        int i = pv.ei[e], j = pv.ej[e];
        float Xi[7], Xj[7], Zm[7], r[6];
        load7(nodes + 7 * (size_t)i, Xi);
        load7(nodes + 7 * (size_t)j, Xj);
        load7(pv.Z + 7 * (size_t)e, Zm);
        if (MODE == 0) {
            float Mm[9], K[9];
            vo_factor(Xi, Xj, Zm, r, Mm, K);
            float* ro = lb.r_vo + 6 * (size_t)e;
#pragma unroll
            for (int k = 0; k < 6; ++k) ro[k] = r[k];
            float* Jo = lb.J_vo + 18 * (size_t)e;
#pragma unroll
            for (int k = 0; k < 9; ++k) { Jo[k] = Mm[k]; Jo[9 + k] = K[k]; }
            // S = w0 J^T J, q = w0 J^T r with J = [[Mm, K],[0, Mm]]  (float64 products)
            double J6[6][6];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    J6[a][b] = Mm[3 * a + b]; J6[a][3 + b] = K[3 * a + b];
                    J6[3 + a][b] = 0.0;       J6[3 + a][3 + b] = Mm[3 * a + b];
                }
            double* So = lb.S_vo + 36 * (size_t)e;
            double* qo = lb.q_vo + 6 * (size_t)e;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                double qa = 0.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) qa += J6[k][a] * (double)r[k];
                qo[a] = pv.w[0] * qa;
#pragma unroll
                for (int b = 0; b < 6; ++b) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < 6; ++k) s += J6[k][a] * J6[k][b];
                    So[6 * a + b] = pv.w[0] * s;
                }
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) lsum += (double)r[k] * (double)r[k];
        } else {
            vo_factor(Xi, Xj, Zm, r, nullptr, nullptr);
#pragma unroll
            for (int k = 0; k < 6; ++k) lsum += (double)r[k] * (double)r[k];
            // quality term with the linearisation-point J and r
            const float* Jo = lb.J_vo + 18 * (size_t)e;
            const float* ro = lb.r_vo + 6 * (size_t)e;
            double d[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) d[k] = D[9 * (size_t)j + k] - D[9 * (size_t)i + k];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                double jt = 0.0, jp = 0.0;
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    jt += (double)Jo[3 * a + b] * d[b] + (double)Jo[9 + 3 * a + b] * d[3 + b];
                    jp += (double)Jo[3 * a + b] * d[3 + b];
                }
                qsum += jt * (2.0 * (double)ro[a] + jt) + jp * (2.0 * (double)ro[3 + a] + jp);
            }
        }
    }
    double tot = block_sum<LIN_THREADS>(lsum, sh);
    if (threadIdx.x == 0) part_out[2 * blk] = tot;
    if (MODE == 1) {
        double tq = block_sum<LIN_THREADS>(qsum, sh2);
        if (threadIdx.x == 0) part_out[2 * blk + 1] = tq;
    } else if (threadIdx.x == 0) {
        part_out[2 * blk + 1] = 0.0;
    }
}

// ---------------------------------------------------------------------------------------------- IMU factors
template <int MODE>
__device__ __forceinline__ void
imu_block(int blk, const LMState* __restrict__ st, const float* __restrict__ nodes0, const float* __restrict__ nodes1,
      const float* __restrict__ vels0, const float* __restrict__ vels1, ProblemView pv, LinBuffers lb,
      const double* __restrict__ D, double* __restrict__ part_out, int force) {
    if (!force) {
        if (!st->active) return;
        if (MODE == 0 && !st->do_lin) return;
    }
    int cur = st->cur;
    const float* nodes = (MODE == 0) ? (cur ? nodes1 : nodes0) : (cur ? nodes0 : nodes1);
    const float* vels = (MODE == 0) ? (cur ? vels1 : vels0) : (cur ? vels0 : vels1);
    __shared__ double sh[LIN_THREADS / 32];
    __shared__ double sh2[LIN_THREADS / 32];
    int i = blk * LIN_THREADS + threadIdx.x;
    double lsum = 0.0, qsum = 0.0;
    bool mine = i < pv.M && (pv.pair_owner == nullptr || pv.pair_owner[i] == pv.part);
    if (mine) {
        float Xa[7], Xb[7], va[3], vb[3], dq[4], r[9];
        load7(nodes + 7 * (size_t)i, Xa);
        load7(nodes + 7 * (size_t)(i + 1), Xb);
#pragma unroll
        for (int k = 0; k < 3; ++k) { va[k] = vels[3 * (size_t)i + k]; vb[k] = vels[3 * (size_t)(i + 1) + k]; }
#pragma unroll
        for (int k = 0; k < 4; ++k) dq[k] = pv.drot[4 * (size_t)i + k];
        float dt = pv.dt[i];
        float Jr[9];
        rot_factor(Xa + 3, Xb + 3, dq, r + 3, MODE == 0 ? Jr : nullptr);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            r[k] = pv.dvel[3 * (size_t)i + k] - (vb[k] - va[k]);                                  // pvgo.py:42
            r[6 + k] = (Xb[k] - Xa[k]) - (va[k] * dt + pv.dtrans[3 * (size_t)i + k]);             // pvgo.py:51
        }
        if (MODE == 0) {
            float* ro = lb.r_imu + 9 * (size_t)i;
            float* Jo = lb.J_rot + 9 * (size_t)i;
#pragma unroll
            for (int k = 0; k < 9; ++k) { ro[k] = r[k]; Jo[k] = Jr[k]; }
#pragma unroll
            for (int k = 0; k < 9; ++k) lsum += (double)r[k] * (double)r[k];
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) lsum += (double)r[k] * (double)r[k];
            const float* ro = lb.r_imu + 9 * (size_t)i;
            const float* Jo = lb.J_rot + 9 * (size_t)i;
            const double* Da = D + 9 * (size_t)i;
            const double* Db = D + 9 * (size_t)(i + 1);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                double j1 = Da[6 + a] - Db[6 + a];                                 // adjvelerr: +I on v_i, -I on v_{i+1}
                double j2 = 0.0;
#pragma unroll
                for (int b = 0; b < 3; ++b) j2 += (double)Jo[3 * a + b] * (Db[3 + b] - Da[3 + b]);
                double j3 = Db[a] - Da[a] - (double)dt * Da[6 + a];                // transvel: [I|0] on tau (A.3 quirk)
                qsum += j1 * (2.0 * (double)ro[a] + j1) + j2 * (2.0 * (double)ro[3 + a] + j2) +
                        j3 * (2.0 * (double)ro[6 + a] + j3);
            }
        }
    }
    double tot = block_sum<LIN_THREADS>(lsum, sh);
    if (threadIdx.x == 0) part_out[2 * blk] = tot;
    if (MODE == 1) {
        double tq = block_sum<LIN_THREADS>(qsum, sh2);
        if (threadIdx.x == 0) part_out[2 * blk + 1] = tq;
    } else if (threadIdx.x == 0) {
        part_out[2 * blk + 1] = 0.0;
    }
}

// ---------------------------------------------------------------------------------------------- reprojection factors
// One warp per consecutive pair i (pvgo.py:53-61 with dense_ba.py:299-305): motion = X_i^-1 X_{i+1} (motion_0 overwritten by
// the constant 0.1, all seven numbers: pvgo.py:57), T = C^-1 motion C, r = point2pixel(P, K, T^-1) - target for the pair's
// points.  Every op on that path is a LieTensor op, so PyPose's Jacobian is the true left-tangent one:
// d r / d delta_i = Pi(p') R_C^T R_j^T [I | -[W]x], W = X_i C P, d r / d delta_{i+1} = -(that); pair 0 is a constant (zero J).
// mode 0 writes r, the unweighted per-pair J^T J / J^T r (float64 sums over the points, fixed order) and the loss partials;
// mode 1 evaluates the trial residuals and the (J D)^T (2 r + J D) term from the stored per-pair sums.
constexpr int RP_PAIRS = LIN_THREADS / 32;
template <int MODE>
__device__ __forceinline__ void
rp_block(int blk, const LMState* __restrict__ st, const float* __restrict__ nodes0, const float* __restrict__ nodes1,
         ProblemView pv, LinBuffers lb, const double* __restrict__ D, double* __restrict__ part_out, int force) {
    if (!force) {
        if (!st->active) return;
        if (MODE == 0 && !st->do_lin) return;
    }
    const int cur = st->cur;
    const float* nodes = (MODE == 0) ? (cur ? nodes1 : nodes0) : (cur ? nodes0 : nodes1);
    __shared__ double shl[RP_PAIRS], shq[RP_PAIRS];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int i = blk * RP_PAIRS + wp;
    const bool mine = i < pv.M && (pv.pair_owner == nullptr || pv.pair_owner[i] == pv.part);
    double lsum = 0.0, qsum = 0.0;
    if (mine) {
        const int np_ = pv.rp_n;
        const float fx = pv.rp_cal[0], fy = pv.rp_cal[1], cx = pv.rp_cal[2], cy = pv.rp_cal[3];
        float C[7], Xi[7], Xj[7], Ci[7], mo[7], T1[7], T[7], Ti[7];
        load7(pv.rp_cal + 4, C);
        load7(nodes + 7 * (size_t)i, Xi);
        load7(nodes + 7 * (size_t)(i + 1), Xj);
        if (i == 0) {
#pragma unroll
            for (int k = 0; k < 7; ++k) mo[k] = 0.1f;                    // pvgo.py:57
        } else {
            float Xii[7];
            se3_inv(Xi, Xii);
            se3_mul(Xii, Xj, mo);                                        // pvgo.py:54-56
        }
        se3_inv(C, Ci);
        se3_mul(Ci, mo, T1);
        se3_mul(T1, C, T);                                               // dense_ba.py:300
        se3_inv(T, Ti);
        float A[9];                                                       // R_C^T R_j^T
        if (MODE == 0) {
            float RC[9], Rj[9];
            q_matrix(C + 3, RC);
            q_matrix(Xj + 3, Rj);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    float v = 0.f;
#pragma unroll
                    for (int k = 0; k < 3; ++k) v += RC[3 * k + a] * Rj[3 * b + k];
                    A[3 * a + b] = v;
                }
        }
        double S[21], q[6];
#pragma unroll
        for (int k = 0; k < 21; ++k) S[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) q[k] = 0.0;
        const float* pts = pv.rp_pts + 3 * (size_t)i * np_;
        const float* tgt = pv.rp_tgt + 2 * (size_t)i * np_;
        for (int p = lane; p < np_; p += 32) {
            const float P[3] = {pts[3 * p], pts[3 * p + 1], pts[3 * p + 2]};
            float pc[3];
            q_rot(Ti + 3, P, pc);
            pc[0] += Ti[0]; pc[1] += Ti[1]; pc[2] += Ti[2];
            const float z = copysignf(fmaxf(fabsf(pc[2]), 1.17549435e-38f), pc[2] >= 0.f ? 1.f : -1.f);   // homo2cart's clamp
            const float r0 = fx * pc[0] / z + cx - tgt[2 * p], r1 = fy * pc[1] / z + cy - tgt[2 * p + 1];
            lsum += (double)r0 * r0 + (double)r1 * r1;
            if (MODE == 0) {
                lb.r_rp[2 * ((size_t)i * np_ + p)] = r0;
                lb.r_rp[2 * ((size_t)i * np_ + p) + 1] = r1;
                if (i > 0) {
                    float B[3], W[3];
                    q_rot(C + 3, P, B);
                    B[0] += C[0]; B[1] += C[1]; B[2] += C[2];
                    q_rot(Xi + 3, B, W);
                    W[0] += Xi[0]; W[1] += Xi[1]; W[2] += Xi[2];
                    // rows of Pi A (2 x 3), then J = (Pi A) [I | -[W]x]
                    const float iz = 1.f / z;
                    const float p0[3] = {fx * iz, 0.f, -fx * pc[0] * iz * iz}, p1[3] = {0.f, fy * iz, -fy * pc[1] * iz * iz};
                    float M0[3], M1[3];
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        M0[b] = p0[0] * A[b] + p0[1] * A[3 + b] + p0[2] * A[6 + b];
                        M1[b] = p1[0] * A[b] + p1[1] * A[3 + b] + p1[2] * A[6 + b];
                    }
                    // row vector m times -[W]x = (m2 W1 - m1 W2, m0 W2 - m2 W0, m1 W0 - m0 W1)
                    double J0[6] = {M0[0], M0[1], M0[2], (double)M0[2] * W[1] - (double)M0[1] * W[2],
                                    (double)M0[0] * W[2] - (double)M0[2] * W[0], (double)M0[1] * W[0] - (double)M0[0] * W[1]};
                    double J1[6] = {M1[0], M1[1], M1[2], (double)M1[2] * W[1] - (double)M1[1] * W[2],
                                    (double)M1[0] * W[2] - (double)M1[2] * W[0], (double)M1[1] * W[0] - (double)M1[0] * W[1]};
                    int k = 0;
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        q[a] += J0[a] * r0 + J1[a] * r1;
#pragma unroll
                        for (int b = a; b < 6; ++b) S[k++] += J0[a] * J0[b] + J1[a] * J1[b];
                    }
                }
            }
        }
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < 21; ++k) S[k] = warp_sum(S[k]);
#pragma unroll
            for (int k = 0; k < 6; ++k) q[k] = warp_sum(q[k]);
            if (lane == 0) {
                double* So = lb.S_rp + 36 * (size_t)i;
                int k = 0;
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int b = a; b < 6; ++b) { So[6 * a + b] = S[k]; So[6 * b + a] = S[k]; ++k; }
#pragma unroll
                for (int a = 0; a < 6; ++a) lb.q_rp[6 * (size_t)i + a] = q[a];
            }
        } else if (lane == 0) {
            // (J D)^T (2 r + J D) summed over the pair's rows = 2 d^T (J^T r) + d^T (J^T J) d, d = D_i - D_{i+1}
            const double* So = lb.S_rp + 36 * (size_t)i;
            const double* qo = lb.q_rp + 6 * (size_t)i;
            double d[6];
#pragma unroll
            for (int a = 0; a < 6; ++a) d[a] = D[9 * (size_t)i + a] - D[9 * (size_t)(i + 1) + a];
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                double sd = 0.0;
#pragma unroll
                for (int b = 0; b < 6; ++b) sd += So[6 * a + b] * d[b];
                qsum += d[a] * (2.0 * qo[a] + sd);
            }
        }
    }
    lsum = warp_sum(lsum);
    if (lane == 0) { shl[wp] = lsum; shq[wp] = qsum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double l = 0.0, qq = 0.0;
#pragma unroll
        for (int k = 0; k < RP_PAIRS; ++k) { l += shl[k]; qq += shq[k]; }
        part_out[2 * blk] = l;
        part_out[2 * blk + 1] = MODE == 1 ? qq : 0.0;
    }
}

// one launch for all factor families: blocks [0, nblk_vo) take VO / loop-closure edges, the next nblk_imu the IMU pairs, the rest
// (if the optional reprojection factor is present) four pairs' reprojection residuals each
template <int MODE>
__global__ void __launch_bounds__(LIN_THREADS)
k_factors(const LMState* __restrict__ st, const float* __restrict__ nodes0, const float* __restrict__ nodes1,
          const float* __restrict__ vels0, const float* __restrict__ vels1, ProblemView pv, LinBuffers lb,
          const double* __restrict__ D, double* __restrict__ part_out, int nblk_vo, int nblk_imu, int force) {
    cudaGridDependencySynchronize();           // PDL: only the launch latency overlaps the previous kernel
    cudaTriggerProgrammaticLaunchCompletion();
    if ((int)blockIdx.x < nblk_vo) vo_block<MODE>(blockIdx.x, st, nodes0, nodes1, pv, lb, D, part_out, force);
    else if ((int)blockIdx.x < nblk_vo + nblk_imu)
        imu_block<MODE>(blockIdx.x - nblk_vo, st, nodes0, nodes1, vels0, vels1, pv, lb, D, part_out + 2 * nblk_vo, force);
    else rp_block<MODE>(blockIdx.x - nblk_vo - nblk_imu, st, nodes0, nodes1, pv, lb, D, part_out + 2 * (nblk_vo + nblk_imu), force);
}

// ---------------------------------------------------------------------------------------------- assembly
struct AsmView {
    const int* node_eoff;    // CSR node -> incident VO edges
    const int* node_edges;
    const int* pair_lo;      // [P]
    const int* pair_hi;
    const int* pair_adj;
    const int* pair_eoff;    // CSR pair -> VO edges
    const int* pair_edges;
    int P;
};

// one warp per node: Hd[n] (9x9, full symmetric) and g[n] (9).  Fixed summation order => deterministic.
__device__ __forceinline__ void
assemble_nodes_block(int blk, const LMState* __restrict__ st, ProblemView pv, LinBuffers lb, AsmView av, double* __restrict__ Hd,
                 double* __restrict__ g, int force) {
    if (!force && !(st->active && st->do_lin)) return;
    int n = (blk * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (n >= pv.N) return;
    int e0 = av.node_eoff[n], e1 = av.node_eoff[n + 1];
    bool has_prev = n > 0 && (pv.pair_owner == nullptr || pv.pair_owner[n - 1] == pv.part);
    bool has_next = n < pv.M && (pv.pair_owner == nullptr || pv.pair_owner[n] == pv.part);
    const float* Jp = lb.J_rot + 9 * (size_t)(n - 1);
    const float* Jn = lb.J_rot + 9 * (size_t)n;
    double dtn = has_next ? (double)pv.dt[n] : 0.0;
    for (int idx = lane; idx < 81; idx += 32) {
        int a = idx / 9, b = idx - 9 * a;
        double v = 0.0;
        if (a < 6 && b < 6) {
            for (int k = e0; k < e1; ++k) {
                int e = av.node_edges[k];
                if (pv.edge_owner != nullptr && pv.edge_owner[e] != pv.part) continue;
                v += lb.S_vo[36 * (size_t)e + 6 * a + b];
            }
        }
        if (a >= 3 && a < 6 && b >= 3 && b < 6) {        // imu rotation: w2 Jrot^T Jrot on phi of both ends
            int aa = a - 3, bb = b - 3;
            if (has_prev) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 3; ++k) s += (double)Jp[3 * k + aa] * (double)Jp[3 * k + bb];
                v += pv.w[2] * s;
            }
            if (has_next) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 3; ++k) s += (double)Jn[3 * k + aa] * (double)Jn[3 * k + bb];
                v += pv.w[2] * s;
            }
        }
        if (a == b) {
            if (a < 3) v += (has_prev ? pv.w[3] : 0.0) + (has_next ? pv.w[3] : 0.0);               // transvel tau
            if (a >= 6) v += (has_prev ? pv.w[1] : 0.0) + (has_next ? pv.w[1] + pv.w[3] * dtn * dtn : 0.0);
        }
        if (has_next && ((a < 3 && b == a + 6) || (b < 3 && a == b + 6))) v += pv.w[3] * dtn;   // tau_i - v_i cross
        if (pv.rp_n > 0 && a < 6 && b < 6) {                                                     // reprojection: +-J on both ends
            if (has_prev) v += pv.w[4] * lb.S_rp[36 * (size_t)(n - 1) + 6 * a + b];
            if (has_next) v += pv.w[4] * lb.S_rp[36 * (size_t)n + 6 * a + b];
        }
        Hd[81 * (size_t)n + idx] = v;
    }
    if (lane < 9) {
        int a = lane;
        double v = 0.0;
        if (a < 6) {
            for (int k = e0; k < e1; ++k) {
                int e = av.node_edges[k];
                if (pv.edge_owner != nullptr && pv.edge_owner[e] != pv.part) continue;
                double q = lb.q_vo[6 * (size_t)e + a];
                v += (pv.ej[e] == n) ? q : -q;
            }
        }
        const float* rp = lb.r_imu + 9 * (size_t)(n - 1);
        const float* rn = lb.r_imu + 9 * (size_t)n;
        if (a < 3) {
            if (has_prev) v += pv.w[3] * (double)rp[6 + a];
            if (has_next) v -= pv.w[3] * (double)rn[6 + a];
        } else if (a < 6) {
            int aa = a - 3;
            if (has_prev) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 3; ++k) s += (double)Jp[3 * k + aa] * (double)rp[3 + k];
                v += pv.w[2] * s;
            }
            if (has_next) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 3; ++k) s += (double)Jn[3 * k + aa] * (double)rn[3 + k];
                v -= pv.w[2] * s;
            }
        } else {
            int aa = a - 6;
            if (has_prev) v -= pv.w[1] * (double)rp[aa];
            if (has_next) v += pv.w[1] * (double)rn[aa] - pv.w[3] * dtn * (double)rn[6 + aa];
        }
        if (pv.rp_n > 0 && a < 6) {                     // J(delta_i) = +J of pair i, J(delta_{i+1}) = -J
            if (has_prev) v -= pv.w[4] * lb.q_rp[6 * (size_t)(n - 1) + a];
            if (has_next) v += pv.w[4] * lb.q_rp[6 * (size_t)n + a];
        }
        g[9 * (size_t)n + a] = v;
    }
}

// one warp per unique pair (lo < hi): Ho[p] = H[lo dofs, hi dofs]
__device__ __forceinline__ void
assemble_pairs_block(int blk, const LMState* __restrict__ st, ProblemView pv, LinBuffers lb, AsmView av, double* __restrict__ Ho,
                 int force) {
    if (!force && !(st->active && st->do_lin)) return;
    int p = (blk * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (p >= av.P) return;
    int lo = av.pair_lo[p];
    int e0 = av.pair_eoff[p], e1 = av.pair_eoff[p + 1];
    bool adj = av.pair_adj[p] && (pv.pair_owner == nullptr || pv.pair_owner[lo] == pv.part);
    const float* Jr = lb.J_rot + 9 * (size_t)lo;
    double dt = adj ? (double)pv.dt[lo] : 0.0;
    // a pair that is not an IMU pair only couples the 6 x 6 pose parts: its velocity rows / columns are structural zeros
    // (zeroed once at creation, never read by the fronts) and are not rewritten every linearisation
    const bool imu_pair = av.pair_adj[p] != 0;
    for (int idx = lane; idx < (imu_pair ? 81 : 36); idx += 32) {
        int a = imu_pair ? idx / 9 : idx / 6, b = imu_pair ? idx - 9 * a : idx - 6 * a;
        double v = 0.0;
        if (a < 6 && b < 6) {
            for (int k = e0; k < e1; ++k) {
                int e = av.pair_edges[k];
                if (pv.edge_owner != nullptr && pv.edge_owner[e] != pv.part) continue;
                v -= lb.S_vo[36 * (size_t)e + 6 * a + b];        // S symmetric: same block for (i,j) and (j,i) edges
            }
        }
        if (adj) {
            if (a >= 3 && a < 6 && b >= 3 && b < 6) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 3; ++k) s += (double)Jr[3 * k + a - 3] * (double)Jr[3 * k + b - 3];
                v -= pv.w[2] * s;
            }
            if (a == b) {
                if (a < 3) v -= pv.w[3];
                if (a >= 6) v -= pv.w[1];
            }
            if (a >= 6 && b < 3 && a - 6 == b) v -= pv.w[3] * dt;   // (v_lo, tau_hi)
            if (pv.rp_n > 0 && adj && a < 6 && b < 6) v -= pv.w[4] * lb.S_rp[36 * (size_t)lo + 6 * a + b];      // owner's share (multi-GPU)
        }
        Ho[81 * (size_t)p + 9 * a + b] = v;
    }
}

// one launch for the block-diagonal and the off-diagonal part of J^T W J (one warp per 9x9 block either way)
__global__ void __launch_bounds__(128)
k_assemble(const LMState* __restrict__ st, ProblemView pv, LinBuffers lb, AsmView av, double* __restrict__ Hd,
           double* __restrict__ Ho, double* __restrict__ g, int nblk_nodes, int force) {
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if ((int)blockIdx.x < nblk_nodes) assemble_nodes_block(blockIdx.x, st, pv, lb, av, Hd, g, force);
    else assemble_pairs_block(blockIdx.x - nblk_nodes, st, pv, lb, av, Ho, force);
}

}  // namespace islam
