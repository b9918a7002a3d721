// Small-graph fast path: the WHOLE `run_pvgo` of one window — linearise, damp, factor, solve, retract, trial loss, trust
// region, accept / roll back, StopOnPlateau, align_to and vo_loss (+ its gradient) — in ONE launch, one CTA per window.
//
// The only workload the reference ships (run_kitti.sh:8, train.py:256-263) calls run_pvgo on 9-pose windows: 81 unknowns.
// On the general path that is ~20 kernel launches per try and a graph replay per try for a problem that fits one SM's
// shared memory many times over; here the normal equations (dense, <= 144 x 144, float64) never leave shared memory and the
// host synchronises once.  `B` windows of identical structure run as a grid of B CTAs (run_pvgo_batch).
//
// Same semantics as the general path, restated from the same sources: residuals / Jacobian blocks of pvgo.py:26-64
// (linearize.cuh: vo_factor, rot_factor), information scalars pvgo.py:125-129, LM.step + TrustRegion + StopOnPlateau as
// configured at pvgo.py:169-180 (lm.cuh: lm_begin_step_b, lm_control), align_to pvgo.py:114-119, vo_loss
// pvgo.py:67-78.  float32 residuals / Jacobians, float64 normal equations and Cholesky, like everywhere else.
#pragma once
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "lie.cuh"
#include "linearize.cuh"
#include "lm.cuh"

namespace islam {

#ifdef ISLAM_PHASE_CLOCKS
__device__ long long g_small_clk[16];
#define SMALL_T(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_small_clk[k] = clock64(); } while (0)
#else
#define SMALL_T(k) do { } while (0)
#endif

constexpr int SM_THREADS = 256;
constexpr int SM_MAX_N = 16;             // poses per window (9 N <= 144 unknowns, dense in shared memory)
constexpr int SM_MAX_E = 128;

struct SmallArgs {
    int nspec;                           // speculative retry slots that fit shared memory (0 / 1: off)
    int N, E, M, B, bw, span;            // span = max |i - j| over the edges (>= 1: the IMU chain); bw = 9 (span + 1) - 1 scalar sub-diagonals
    const int* links;                    // [E, 2] int32, shared by all windows
    const float *nodes0, *vels0, *Z, *drot, *dtrans, *dvel, *dt;          // batched, window-major
    double w[4];
    islam_lm_params prm;
    float *out_nodes, *out_vels;         // aligned to the window's first initial pose (pvgo.py:195)
    islam_lm_state* out_state;
    const float* voP;                    // [B, E, 7] motions for the outer loss (nullable: Z itself)
    float *tl, *rl, *gt, *gr;            // vo_loss outputs (nullable)
};

__host__ __device__ inline size_t small_smem_bytes(int N, int E) {
    const int n = 9 * N, ld = n | 1, M = N - 1;
    size_t d = (size_t)ld * n + 7 * (size_t)n + 16;                          // A, diag0, dinv, g, y, D, diagW (+ slack)
    size_t f = 2 * (size_t)N * 10 + (size_t)E * (7 + 6 + 18) + (size_t)M * (4 + 3 + 3 + 1 + 9 + 9);
    return d * sizeof(double) + (f + 2 * (size_t)E + 8) * sizeof(float);
}


// ---- pieces of one LM try that a single warp can run on its own (the sequential path and the speculative retries share
// them, so both produce bit-identical numbers) ---------------------------------------------------------------------------

// Block-banded Cholesky with the right-hand side riding along, by ONE warp (scheme of front4.cuh: every lane holds the whole
// 9 x 9 diagonal block in registers and factors it redundantly, lane l owns row l below the block, the lane after the last
// row the right-hand side).  W(row, col) = W[col + row * ldw] is the working copy of the strictly lower band (row i of L),
// diagW the working diagonal, yv the right-hand side; on return W / dinv / yv hold L, 1 / L_jj and the forward-substituted rhs.
__device__ __forceinline__ bool small_factor_warp(double* __restrict__ W, int ldw, double* __restrict__ diagW,
                                                  double* __restrict__ yv, double* __restrict__ dinv, int n, int bwn, int lane) {
    bool ok = true;
    for (int c0 = 0; c0 < n; c0 += 9) {
        const int nbelow = min(bwn, n - c0 - 9);
        const bool is_row = lane < nbelow, is_rhs = lane == nbelow;
        const int i = c0 + 9 + lane;
        double D[9][9], isv[9], r[9];
#pragma unroll
        for (int p = 0; p < 9; ++p) {
            D[p][p] = diagW[c0 + p];
#pragma unroll
            for (int q = 0; q < p; ++q) D[p][q] = W[(c0 + q) + (c0 + p) * ldw];
        }
        double* rowp = is_row ? W + c0 + i * ldw : yv + c0;       // this lane's nine entries of the block column
#pragma unroll
        for (int q = 0; q < 9; ++q) r[q] = (is_row || is_rhs) ? rowp[q] : 0.0;
        F4Diag<0>::run(D, isv, ok);
        f4_row_solve(r, D, isv);
        __syncwarp();                                   // every lane has read the block before any lane overwrites it with L
#pragma unroll
        for (int p = 0; p < 9; ++p) {
            dinv[c0 + p] = isv[p];
#pragma unroll
            for (int q = 0; q < p; ++q) W[(c0 + q) + (c0 + p) * ldw] = D[p][q];
        }
        if (is_row || is_rhs) {
#pragma unroll
            for (int q = 0; q < 9; ++q) rowp[q] = r[q];
        }
        __syncwarp();
        // trailing update inside the band: W[i][j'] -= sum_q L[i][q] L[j'][q] for the rows j' <= i below the block
        // (nine rows j' at a time, all accumulators independent: the loads are issued together, not one
        // dependent load -> fma chain per entry)
        for (int u0 = 0; u0 < nbelow; u0 += 9) {
            double s_[9];
#pragma unroll
            for (int u = 0; u < 9; ++u) s_[u] = 0.0;
#pragma unroll
            for (int q = 0; q < 9; ++q) {
#pragma unroll
                for (int u = 0; u < 9; ++u) {
                    const int uu = u0 + u < nbelow ? u0 + u : u0;
                    s_[u] = fma(r[q], W[c0 + q + (c0 + 9 + uu) * ldw], s_[u]);       // L[j'][c0 + q], broadcast
                }
            }
#pragma unroll
            for (int u = 0; u < 9; ++u) {
                const int uu = u0 + u;
                if (uu >= nbelow) continue;
                if (is_rhs) yv[c0 + 9 + uu] -= s_[u];
                else if (is_row && lane == uu) diagW[i] -= s_[u];
                else if (is_row && lane > uu) W[(c0 + 9 + uu) + i * ldw] -= s_[u];
            }
        }
        __syncwarp();
    }
    return ok;
}

// L^T D = y by one warp (Dv starts as y)
__device__ __forceinline__ void small_backsolve_warp(const double* __restrict__ W, int ldw, const double* __restrict__ dinv,
                                                     double* __restrict__ Dv, int n, int bw, int lane) {
    int offj = (n - 1) * ldw;
    for (int j = n - 1; j >= 0; --j, offj -= ldw) {
        const double dj = Dv[j] * dinv[j];
        __syncwarp();
        if (lane == 0) Dv[j] = dj;
        const int i = j - 1 - lane;
        if (i >= 0 && lane < bw) Dv[i] = fma(-W[offj + i], dj, Dv[i]);          // L[j][i]
        __syncwarp();
    }
}

// node nd of the trial state: nodes <- Exp(d[:6]) nodes ; vels <- vels + d[6:9]   (LieTensor.add_, A.1)
__device__ __forceinline__ void small_retract_node(int nd, const double* __restrict__ Dv, const float* __restrict__ Xc,
                                                   const float* __restrict__ Vc, float* __restrict__ Xt, float* __restrict__ Vt) {
    double xi[6], Xd[7], Ed[7], Od[7];
#pragma unroll
    for (int k = 0; k < 6; ++k) xi[k] = Dv[9 * nd + k];
#pragma unroll
    for (int k = 0; k < 7; ++k) Xd[k] = (double)Xc[7 * nd + k];
    se3_exp(xi, Ed);
    se3_mul(Ed, Xd, Od);
#pragma unroll
    for (int k = 0; k < 7; ++k) Xt[7 * nd + k] = (float)Od[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) Vt[3 * nd + k] = (float)((double)Vc[3 * nd + k] + Dv[9 * nd + 6 + k]);
}

struct SmallFactors {                    // the staged window (shared memory)
    int E;
    const int *ei, *ej;
    const float *Zs, *r_vo, *J_vo, *sdrot, *sdtr, *sdv, *sdt, *r_imu, *J_rot;
};
// factor f at the trial state: its share of sum r^2 and of the unweighted (J D)^T (2 r + J D) of TrustRegion.update (A.4)
__device__ __forceinline__ void small_trial_factor(int f, const SmallFactors& F, const float* __restrict__ Xt,
                                                   const float* __restrict__ Vt, const double* __restrict__ Dv, double& lsum,
                                                   double& qsum) {
    if (f < F.E) {
        const int i = F.ei[f], j = F.ej[f];
        float r[6];
        vo_factor(Xt + 7 * i, Xt + 7 * j, F.Zs + 7 * f, r, nullptr, nullptr);
#pragma unroll
        for (int k = 0; k < 6; ++k) lsum += (double)r[k] * r[k];
        const float* Jo = F.J_vo + 18 * f;
        const float* ro = F.r_vo + 6 * f;
        double d[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) d[k] = Dv[9 * j + k] - Dv[9 * i + k];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            double jt = 0.0, jp = 0.0;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                jt += (double)Jo[3 * p + q] * d[q] + (double)Jo[9 + 3 * p + q] * d[3 + q];
                jp += (double)Jo[3 * p + q] * d[3 + q];
            }
            qsum += jt * (2.0 * (double)ro[p] + jt) + jp * (2.0 * (double)ro[3 + p] + jp);
        }
    } else {
        const int i = f - F.E;
        const float* Xa = Xt + 7 * i;
        const float* Xb = Xa + 7;
        float r[9];
        rot_factor(Xa + 3, Xb + 3, F.sdrot + 4 * i, r + 3, nullptr);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            r[k] = F.sdv[3 * i + k] - (Vt[3 * (i + 1) + k] - Vt[3 * i + k]);
            r[6 + k] = (Xb[k] - Xa[k]) - (Vt[3 * i + k] * F.sdt[i] + F.sdtr[3 * i + k]);
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) lsum += (double)r[k] * r[k];
        const float* ro = F.r_imu + 9 * i;
        const float* Jo = F.J_rot + 9 * i;
        const double* Da = Dv + 9 * i;
        const double* Db = Da + 9;
        const double dt_ = (double)F.sdt[i];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const double j1 = Da[6 + p] - Db[6 + p];
            double j2 = 0.0;
#pragma unroll
            for (int q = 0; q < 3; ++q) j2 += (double)Jo[3 * p + q] * (Db[3 + q] - Da[3 + q]);
            const double j3 = Db[p] - Da[p] - dt_ * Da[6 + p];
            qsum += j1 * (2.0 * (double)ro[p] + j1) + j2 * (2.0 * (double)ro[3 + p] + j2) + j3 * (2.0 * (double)ro[6 + p] + j3);
        }
    }
}

// Speculative retries.  When a try is rejected PyPose re-damps the SAME normal equations and tries again, up to 16 times
// (the shipped window does exactly that at convergence: 2 steps = 18 tries).  The damping of retry k+1 depends on retry k
// only through the quality CLASS of TrustRegion.update, so the next SM_SPEC dampings are predicted (same class as the try
// just rejected), factored, solved and evaluated at once, one warp each on its own band copy, and the controller then
// consumes the results in order — the real lm_control runs for every consumed try and the batch is cut short the moment
// its state departs from the prediction (or a try is accepted), so the sequence of decisions is exactly the sequential one.
constexpr int SM_SPEC = SM_THREADS / 32;
__host__ __device__ inline size_t small_slot_doubles(int N, int span) {
    const int n = 9 * N, S = 9 * span + 9;
    return ((size_t)n * S + 4 * (size_t)n + 5 * (size_t)N + 2) & ~(size_t)1;        // band, diagW, yv, dinv, Dv, trial state (10 N floats)
}

__global__ void __launch_bounds__(SM_THREADS, 1) k_lm_small(SmallArgs a) {
    extern __shared__ double smem_d[];
    __shared__ LMState st;
    __shared__ islam_lm_params prm;
    __shared__ double red[SM_THREADS / 32], red2[SM_THREADS / 32], sums[4];
    __shared__ int fail;
    const int tid = threadIdx.x, NT = SM_THREADS, win = blockIdx.x;
    const int N = a.N, E = a.E, M = a.M, n = 9 * N, ld = n | 1, bw = a.bw;
    double* A = smem_d;                       // lower triangle: J^T W J;  strict upper: L^T (row i of L in column i)
    double* diag0 = A + (size_t)ld * n;       // clamped diagonal of J^T W J
    double* dinv = diag0 + n;                 // 1 / L_jj
    double* gv = dinv + n;                    // J^T W r
    double* yv = gv + n;                      // forward-substituted right-hand side
    double* Dv = yv + n;                      // step
    double* diagW = Dv + n;                   // working diagonal of the factorisation
    float* fb = reinterpret_cast<float*>(diagW + n + 2);
    float* X[2] = {fb, fb + 7 * N};
    float* V[2] = {fb + 14 * N, fb + 17 * N};
    float* Zs = fb + 20 * N;
    float* r_vo = Zs + 7 * E;
    float* J_vo = r_vo + 6 * E;
    float* sdrot = J_vo + 18 * E;
    float* sdtr = sdrot + 4 * M;
    float* sdv = sdtr + 3 * M;
    float* sdt = sdv + 3 * M;
    float* r_imu = sdt + M;
    float* J_rot = r_imu + 9 * M;
    int* ei = reinterpret_cast<int*>(J_rot + 9 * M);
    int* ej = ei + E;
    // ---- stage the window -------------------------------------------------------------------------------------------
    for (int k = tid; k < 7 * N; k += NT) X[0][k] = a.nodes0[(size_t)win * 7 * N + k];
    for (int k = tid; k < 3 * N; k += NT) V[0][k] = a.vels0[(size_t)win * 3 * N + k];
    for (int k = tid; k < 7 * E; k += NT) Zs[k] = a.Z[(size_t)win * 7 * E + k];
    for (int k = tid; k < 4 * M; k += NT) sdrot[k] = a.drot[(size_t)win * 4 * M + k];
    for (int k = tid; k < 3 * M; k += NT) { sdtr[k] = a.dtrans[(size_t)win * 3 * M + k]; sdv[k] = a.dvel[(size_t)win * 3 * M + k]; }
    for (int k = tid; k < M; k += NT) sdt[k] = a.dt[(size_t)win * M + k];
    for (int k = tid; k < E; k += NT) { ei[k] = a.links[2 * k]; ej[k] = a.links[2 * k + 1]; }
    if (tid == 0) {
        prm = a.prm;
        memset(&st, 0, sizeof(st));
        st.damping = 1.0 / prm.radius; st.radius = prm.radius; st.down = prm.down; st.diag_scale = 1.0;
        st.need_linearize = 1; st.continual = 1;
    }
    __syncthreads();
    const double w0 = a.w[0], w1 = a.w[1], w2 = a.w[2], w3 = a.w[3];
    const int max_tries = (prm.max_steps > 0 ? prm.max_steps : 1) * (prm.reject + 1) + 1;
    const SmallFactors FW{E, ei, ej, Zs, r_vo, J_vo, sdrot, sdtr, sdv, sdt, r_imu, J_rot};
    // speculative retry slots behind the window (8-byte aligned)
    const bool spec_ok = a.nspec > 1 && a.span <= 2;
    double* slots = smem_d + ((small_smem_bytes(N, E) + 7) >> 3);
    const size_t slot_doubles = small_slot_doubles(N, a.span);
    __shared__ double sp_scale[SM_SPEC], sp_damp[SM_SPEC], sp_S[SM_SPEC], sp_Q[SM_SPEC];
    __shared__ int sp_fail[SM_SPEC], sp_n, sp_take;

    for (int it = 0; it < max_tries; ++it) {
        if (!st.continual) break;                       // (uniform: written before the barrier that ends every try)
        __syncthreads();
        if (spec_ok && !st.need_linearize) {
            // ---- speculative retries (see SM_SPEC above): the normal equations are those of the rejected try --------------
            const int cur = st.cur;
            if (tid == 0) {
                double damping = st.damping, down = st.down, scale = st.diag_scale;
                const int left = prm.reject + 1 - st.reject_count;      // tries until the step accepts whatever comes
                const int m = min(a.nspec, left > 1 ? left : 1);
                // prediction: the retries fall into the same quality class as the try that was just rejected.  (At the noise
                // floor — the shipped window's 16-reject step — that is quality ~ +1: the unweighted model predicts the tiny
                // increase it then gets, so TrustRegion GROWS the radius on every rejected try.)
                const int cls = st.quality > prm.high ? 2 : (st.quality > prm.low ? 1 : 0);
                for (int k = 0; k < m; ++k) {
                    sp_damp[k] = damping;
                    scale *= (1.0 + damping);
                    sp_scale[k] = scale;
                    double radius = 1.0 / damping;                       // TrustRegion.update as in lm_control
                    if (cls == 2) { radius *= prm.up; down = prm.down; }
                    else if (cls == 1) { down = prm.down; }
                    else { radius *= down; down *= prm.factor; }
                    down = fmax(prm.tr_min, fmin(down, prm.tr_max));
                    radius = fmax(prm.tr_min, fmin(radius, prm.tr_max));
                    damping = 1.0 / radius;
                }
                sp_n = m; sp_take = -1;
            }
            __syncthreads();
            const int w = tid >> 5, lane = tid & 31;
            const int S = 9 * a.span + 9, ldw = S - 1, bwn = 9 * a.span;
            double* sW = slots + (size_t)w * slot_doubles;               // band rows at stride S: W(row, col) = Wp[col + row * ldw]
            double* Wp = sW + (S - 1);
            double* sdiag = sW + (size_t)n * S;
            double* sy = sdiag + n;
            double* sdinv = sy + n;
            double* sD = sdinv + n;
            float* sX = reinterpret_cast<float*>(sD + n);
            float* sV = sX + 7 * N;
            if (w < sp_n) {
                for (int idx = lane; idx < n * ldw; idx += 32) {
                    const int i = idx / ldw, j = i - 1 - (idx - i * ldw);
                    if (j >= 0) Wp[j + i * ldw] = A[i + (size_t)j * ld];
                }
                const double scale = sp_scale[w];
                for (int k = lane; k < n; k += 32) { sdiag[k] = diag0[k] * scale; sy[k] = -gv[k]; }
                __syncwarp();
                const bool ok = small_factor_warp(Wp, ldw, sdiag, sy, sdinv, n, bwn, lane);
                for (int k = lane; k < n; k += 32) sD[k] = sy[k];
                __syncwarp();
                small_backsolve_warp(Wp, ldw, sdinv, sD, n, bw, lane);
                __syncwarp();
                for (int nd = lane; nd < N; nd += 32) small_retract_node(nd, sD, X[cur], V[cur], sX, sV);
                __syncwarp();
                // the two sums in the reduction tree of block_sum<SM_THREADS> (factor f in thread f): same bits
                double tl_ = 0.0, tq_ = 0.0;
                for (int c = 0; c < SM_THREADS / 32; ++c) {
                    double l = 0.0, q = 0.0;
                    const int f = 32 * c + lane;
                    if (f < E + M) small_trial_factor(f, FW, sX, sV, sD, l, q);
                    l = warp_sum(l); q = warp_sum(q);
                    l = __shfl_sync(0xffffffffu, l, 0); q = __shfl_sync(0xffffffffu, q, 0);
                    if (lane == c) { tl_ = l; tq_ = q; }
                }
                tl_ = warp_sum(tl_); tq_ = warp_sum(tq_);
                if (lane == 0) { sp_S[w] = tl_; sp_Q[w] = tq_; sp_fail[w] = ok ? 0 : 1; }
            }
            __syncthreads();
            if (tid == 0) {
                for (int k = 0; k < sp_n; ++k) {
                    if (st.damping != sp_damp[k]) break;                 // the controller left the predicted schedule
                    st.active = 1; st.do_lin = 0; st.chol_fail = sp_fail[k]; st.tries_total += 1;
                    st.diag_scale *= (1.0 + st.damping);
                    const int before = st.cur;
                    lm_control(&st, &prm, sp_fail[k] ? 0.0 : sp_S[k], sp_fail[k] ? 0.0 : sp_Q[k]);
                    if (st.cur != before) sp_take = k;                   // accepted: slot k holds the new state
                    if (st.need_linearize || !st.continual) break;       // the step is over (accepted or abandoned)
                }
            }
            __syncthreads();
            if (sp_take >= 0) {
                const float* tX = reinterpret_cast<const float*>(slots + (size_t)sp_take * slot_doubles + (size_t)n * S + 4 * n);
                const float* tV = tX + 7 * N;
                for (int k = tid; k < 7 * N; k += NT) X[st.cur][k] = tX[k];
                for (int k = tid; k < 3 * N; k += NT) V[st.cur][k] = tV[k];
            }
            __syncthreads();
            continue;
        }
        if (tid == 0) { st.active = 1; st.do_lin = st.need_linearize; st.chol_fail = 0; st.tries_total += 1; fail = 0; }
        __syncthreads();
        const int cur = st.cur, do_lin = st.do_lin;
        SMALL_T(0);
        if (do_lin) {
            // ---- kernel family 1: residuals and Jacobian blocks at the current state (pvgo.py:26-64, A.3) --------------
            double lsum = 0.0;
            for (int f = tid; f < E + M; f += NT) {
                if (f < E) {
                    float r[6], Mm[9], K[9];
                    vo_factor(X[cur] + 7 * ei[f], X[cur] + 7 * ej[f], Zs + 7 * f, r, Mm, K);
#pragma unroll
                    for (int k = 0; k < 6; ++k) { r_vo[6 * f + k] = r[k]; lsum += (double)r[k] * r[k]; }
#pragma unroll
                    for (int k = 0; k < 9; ++k) { J_vo[18 * f + k] = Mm[k]; J_vo[18 * f + 9 + k] = K[k]; }
                } else {
                    const int i = f - E;
                    const float* Xa = X[cur] + 7 * i;
                    const float* Xb = Xa + 7;
                    float r[9], Jr[9];
                    rot_factor(Xa + 3, Xb + 3, sdrot + 4 * i, r + 3, Jr);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        r[k] = sdv[3 * i + k] - (V[cur][3 * (i + 1) + k] - V[cur][3 * i + k]);                    // pvgo.py:42
                        r[6 + k] = (Xb[k] - Xa[k]) - (V[cur][3 * i + k] * sdt[i] + sdtr[3 * i + k]);              // pvgo.py:51
                    }
#pragma unroll
                    for (int k = 0; k < 9; ++k) { r_imu[9 * i + k] = r[k]; J_rot[9 * i + k] = Jr[k]; lsum += (double)r[k] * r[k]; }
                }
            }
            const double tot = block_sum<SM_THREADS>(lsum, red);
            if (tid == 0) sums[0] = tot;
            for (int k = tid; k < ld * n; k += NT) A[k] = 0.0;
            __syncthreads();
            SMALL_T(1);
            // ---- assembly: one thread per 9x9 block (bi >= bj) of J^T W J, every factor that touches it, fixed order -----
            const int nblk = N * (N + 1) / 2;
            for (int blk = tid; blk < nblk; blk += NT) {
                int bi = 0;
                while ((bi + 1) * (bi + 2) / 2 <= blk) ++bi;
                const int bj = blk - bi * (bi + 1) / 2;
                double* Ab = A + 9 * bi + (size_t)(9 * bj) * ld;
                auto add = [&](int r_, int c_, double v) { if (bi != bj || r_ >= c_) Ab[r_ + (size_t)c_ * ld] += v; };
                for (int e = 0; e < E; ++e) {
                    const int i = ei[e], j = ej[e];
                    const bool diag = bi == bj && (i == bi || j == bi);
                    const bool off = bi != bj && ((i == bi && j == bj) || (i == bj && j == bi));
                    if (!diag && !off) continue;
                    // J = [[Mm, K], [0, Mm]] (d r / d delta_j; d r / d delta_i = -J):  S = w0 J^T J
                    const float* Jo = J_vo + 18 * e;
                    double J6[6][6];
#pragma unroll
                    for (int p = 0; p < 3; ++p)
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            J6[p][q] = Jo[3 * p + q]; J6[p][3 + q] = Jo[9 + 3 * p + q];
                            J6[3 + p][q] = 0.0;       J6[3 + p][3 + q] = Jo[3 * p + q];
                        }
                    const double sgn = diag ? w0 : -w0;
#pragma unroll
                    for (int p = 0; p < 6; ++p)
#pragma unroll
                        for (int q = 0; q < 6; ++q) {
                            double s_ = 0.0;
#pragma unroll
                            for (int k = 0; k < 6; ++k) s_ += J6[k][p] * J6[k][q];
                            add(p, q, sgn * s_);
                        }
                }
                if (bi == bj) {
                    const int nd = bi;
                    const bool has_prev = nd > 0, has_next = nd < M;
                    const float* Jp = J_rot + 9 * (nd - 1);
                    const float* Jn = J_rot + 9 * nd;
                    const double dtn = has_next ? (double)sdt[nd] : 0.0;
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            double s_ = 0.0;
                            if (has_prev) for (int k = 0; k < 3; ++k) s_ += (double)Jp[3 * k + p] * (double)Jp[3 * k + q];
                            if (has_next) for (int k = 0; k < 3; ++k) s_ += (double)Jn[3 * k + p] * (double)Jn[3 * k + q];
                            add(3 + p, 3 + q, w2 * s_);                                                  // imu rotation
                        }
                        add(p, p, (has_prev ? w3 : 0.0) + (has_next ? w3 : 0.0));                          // trans-vel on tau
                        add(6 + p, 6 + p, (has_prev ? w1 : 0.0) + (has_next ? w1 + w3 * dtn * dtn : 0.0));
                        if (has_next) add(6 + p, p, w3 * dtn);                                             // (v_n, tau_n)
                    }
                    // g = J^T W r of this node
                    const float* rp = r_imu + 9 * (nd - 1);
                    const float* rn = r_imu + 9 * nd;
                    double gl[9];
#pragma unroll
                    for (int p = 0; p < 9; ++p) gl[p] = 0.0;
                    for (int e = 0; e < E; ++e) {
                        const int i = ei[e], j = ej[e];
                        if (i != nd && j != nd) continue;
                        const float* Jo = J_vo + 18 * e;
                        const float* ro = r_vo + 6 * e;
                        const double sg = (j == nd) ? w0 : -w0;
#pragma unroll
                        for (int p = 0; p < 3; ++p) {
                            double qt = 0.0, qp = 0.0;
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                qt += (double)Jo[3 * k + p] * ro[k];
                                qp += (double)Jo[9 + 3 * k + p] * ro[k] + (double)Jo[3 * k + p] * ro[3 + k];
                            }
                            gl[p] += sg * qt; gl[3 + p] += sg * qp;
                        }
                    }
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        if (has_prev) {
                            double s_ = 0.0;
                            for (int k = 0; k < 3; ++k) s_ += (double)Jp[3 * k + p] * (double)rp[3 + k];
                            gl[p] += w3 * (double)rp[6 + p]; gl[3 + p] += w2 * s_; gl[6 + p] -= w1 * (double)rp[p];
                        }
                        if (has_next) {
                            double s_ = 0.0;
                            for (int k = 0; k < 3; ++k) s_ += (double)Jn[3 * k + p] * (double)rn[3 + k];
                            gl[p] -= w3 * (double)rn[6 + p]; gl[3 + p] -= w2 * s_;
                            gl[6 + p] += w1 * (double)rn[p] - w3 * dtn * (double)rn[6 + p];
                        }
                    }
#pragma unroll
                    for (int p = 0; p < 9; ++p) gv[9 * nd + p] = gl[p];
                } else if (bi == bj + 1) {
                    // adjacent pair lo = bj, hi = bi: rows hi, columns lo
                    const int lo = bj;
                    const float* Jr = J_rot + 9 * lo;
                    const double dt_ = (double)sdt[lo];
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            double s_ = 0.0;
                            for (int k = 0; k < 3; ++k) s_ += (double)Jr[3 * k + p] * (double)Jr[3 * k + q];
                            add(3 + p, 3 + q, -w2 * s_);
                        }
                        add(p, p, -w3);
                        add(6 + p, 6 + p, -w1);
                        add(p, 6 + p, -w3 * dt_);                                                          // (tau_hi, v_lo)
                    }
                }
            }
            __syncthreads();
            for (int k = tid; k < n; k += NT) diag0[k] = fmin(fmax(A[k + (size_t)k * ld], prm.lm_min), prm.lm_max);   // clamp_ (A.4)
            if (tid == 0) { st.reject_count = 0; st.diag_scale = 1.0; }
            __syncthreads();
        }
        if (tid == 0) {
            st.diag_scale *= (1.0 + st.damping);                                 // A.diag += A.diag * damping, cumulative
            lm_begin_step_b(&st, sums);
        }
        __syncthreads();
        SMALL_T(2);
        // ---- kernel family 2: block-banded Cholesky (9 x 9 blocks) with the right-hand side -g riding along as one more row ---
        // Chain windows (edges between neighbours up to two poses apart: at most 18 rows below a diagonal block) are factored
        // by ONE warp with the register-resident scheme of front4.cuh: every lane holds the whole diagonal block and factors it
        // redundantly (no shuffle, no load on the serial path), lane l owns row l below the block (the last lane the right-hand
        // side) and forward-substitutes it against the register copy, then the trailing update inside the band.  J^T W J stays
        // intact in the lower triangle (a rejected try re-damps and re-factors it); the working copy and then L live in the
        // strict upper triangle, transposed (row i of L in column i), the working diagonal in diagW.
        const bool warp_path = a.span <= 2;
        {
            const double scale = st.diag_scale;
            if (warp_path) {
                const int bwn = 9 * a.span;                                   // rows below a diagonal block inside the band
                for (int idx = tid; idx < n * (bwn + 9); idx += NT) {          // working copy of the lower band, transposed
                    const int i = idx / (bwn + 9), j = i - (idx - i * (bwn + 9)) - 1;           // j = i-1 .. i-(bwn+9)
                    if (j >= 0) A[j + i * ld] = A[i + j * ld];
                }
                for (int k = tid; k < n; k += NT) { diagW[k] = diag0[k] * scale; yv[k] = -gv[k]; }
                __syncthreads();
                if (tid < 32) {
                    const bool ok = small_factor_warp(A, ld, diagW, yv, dinv, n, bwn, tid);
                    if (!ok) fail = 1;
                }
                __syncthreads();
            } else {
                for (int j = 0; j < n; ++j) {
                    // rows i = j .. min(n-1, j + bw), plus the right-hand-side "row" n
                    const int nrows = min(n - 1, j + bw) - j + 1;
                    double v = 0.0;
                    int i = -1;
                    if (tid <= nrows) {
                        i = tid < nrows ? j + tid : n;
                        const int ks = (i < n) ? max(0, i - bw) : max(0, j - bw);
                        const double* Li = (i < n) ? A + (size_t)i * ld : yv;           // row i of L (or y) for columns k
                        const double* Lj = A + (size_t)j * ld;
                        double s_ = (i == j) ? diag0[j] * scale : (i < n ? A[i + (size_t)j * ld] : -gv[j]);
                        for (int k = ks; k < j; ++k) s_ -= Li[k] * Lj[k];
                        v = s_;
                        if (i == j) {
                            if (!(v > 0.0) || !(v < 1e300)) { fail = 1; v = 1.0; }
                            dinv[j] = rsqrt(v);
                        }
                    }
                    __syncthreads();
                    if (i > j) {
                        const double l = v * dinv[j];
                        if (i < n) A[j + (size_t)i * ld] = l; else yv[j] = l;
                    }
                    __syncthreads();
                }
            }
        }
        if (fail) {
            if (tid == 0) { st.chol_fail = 1; lm_control(&st, &prm, 0.0, 0.0); }      // "Linear solver failed": step abandoned
            __syncthreads();
            continue;
        }
        SMALL_T(3);
        // ---- back-substitution L^T D = y --------------------------------------------------------------------------------
        for (int k = tid; k < n; k += NT) Dv[k] = yv[k];
        __syncthreads();
        if (warp_path) {
            if (tid < 32) small_backsolve_warp(A, ld, dinv, Dv, n, bw, tid);
            __syncthreads();
        } else {
            for (int j = n - 1; j >= 0; --j) {
                const double dj = Dv[j] * dinv[j];
                __syncthreads();
                if (tid == 0) Dv[j] = dj;
                const int i = j - 1 - tid;
                if (i >= 0 && i >= j - bw) Dv[i] -= A[i + (size_t)j * ld] * dj;            // L[j][i]
                __syncthreads();
            }
        }
        SMALL_T(4);
        // ---- retract to the trial state (LieTensor.add_, A.1) --------------------------------------------------------------
        for (int nd = tid; nd < N; nd += NT) small_retract_node(nd, Dv, X[cur], V[cur], X[cur ^ 1], V[cur ^ 1]);
        __syncthreads();
        SMALL_T(5);
        // ---- trial residuals and the unweighted (J D)^T (2 r + J D) of TrustRegion.update (A.4) ---------------------------
        {
            double lsum = 0.0, qsum = 0.0;
            const float* Xt = X[cur ^ 1];
            const float* Vt = V[cur ^ 1];
            for (int f = tid; f < E + M; f += NT) small_trial_factor(f, FW, Xt, Vt, Dv, lsum, qsum);
            const double S = block_sum<SM_THREADS>(lsum, red);
            const double Q = block_sum<SM_THREADS>(qsum, red2);
            if (tid == 0) lm_control(&st, &prm, S, Q);
        }
        __syncthreads();
        SMALL_T(6);
    }
    __syncthreads();
    // ---- outputs: align_to the first initial pose (pvgo.py:114-119,195), LM state, outer loss (pvgo.py:67-78) ---------------
    const int cur = st.cur;
    const float* Xf = X[cur];
    const float* Vf = V[cur];
    for (int nd = tid; nd < N; nd += NT) {
        float Tg[7], X0i[7], T[7], O[7], q0i[4], qr[4], vo[3];
        load7(a.nodes0 + (size_t)win * 7 * N, Tg);
        se3_inv(Xf, X0i);
        se3_mul(Tg, X0i, T);
        se3_mul(T, Xf + 7 * nd, O);
#pragma unroll
        for (int k = 0; k < 7; ++k) a.out_nodes[((size_t)win * N + nd) * 7 + k] = O[k];
        q_inv(Xf + 3, q0i);
        q_mul(Tg + 3, q0i, qr);
        q_rot(qr, Vf + 3 * nd, vo);
#pragma unroll
        for (int k = 0; k < 3; ++k) a.out_vels[((size_t)win * N + nd) * 3 + k] = vo[k];
    }
    if (tid == 0) a.out_state[win] = st;
    if (a.tl != nullptr) {
        for (int e = tid; e < E; e += NT) {
            const float* Pm = (a.voP ? a.voP : a.Z) + ((size_t)win * E + e) * 7;
            float Idn[7] = {0, 0, 0, 0, 0, 0, 1}, C[7], Xii[7], r[6], Mm[9], K[9], Pl[7];
            load7(Pm, Pl);
            se3_inv(Xf + 7 * ei[e], Xii);
            se3_mul(Xii, Xf + 7 * ej[e], C);
            vo_factor(Idn, C, Pl, r, Mm, K);                                   // e = Log(P^-1 n1^-1 n2), J_P = -Jl^-1(e) Ad(P^-1)
            const size_t o = (size_t)win * E + e;
            a.tl[o] = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
            a.rl[o] = r[3] * r[3] + r[4] * r[4] + r[5] * r[5];
            if (a.gt != nullptr) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    float x = 0.f, y = 0.f, z = 0.f;
#pragma unroll
                    for (int q = 0; q < 3; ++q) { x += Mm[3 * q + k] * r[q]; y += K[3 * q + k] * r[q]; z += Mm[3 * q + k] * r[3 + q]; }
                    a.gt[6 * o + k] = -2.f * x; a.gt[6 * o + 3 + k] = -2.f * y;
                    a.gr[6 * o + k] = 0.f;      a.gr[6 * o + 3 + k] = -2.f * z;
                }
            }
        }
    }
}

}  // namespace islam

extern "C" int islam_pvgo_small_supported(int32_t N, int32_t E) { return N >= 2 && N <= islam::SM_MAX_N && E >= 0 && E <= islam::SM_MAX_E; }

extern "C" int islam_pvgo_small_run(int32_t B, int32_t N, int32_t E, const int32_t* links_dev, const int64_t* links_host,
                                    const float* nodes0, const float* vels0, const float* Z, const float* drot, const float* dtrans,
                                    const float* dvel, const float* dt, const double w[4], const islam_lm_params* prm,
                                    float* out_nodes, float* out_vels, islam_lm_state* out_state, const float* voP, float* tl,
                                    float* rl, float* gt, float* gr, void* stream) {
    if (B <= 0 || !islam_pvgo_small_supported(N, E) || !links_dev || (E > 0 && !links_host) || !nodes0 || !vels0 || !drot ||
        !dtrans || !dvel || !dt || !w || !prm || !out_nodes || !out_vels || !out_state || (E > 0 && !Z) ||
        ((tl == nullptr) != (rl == nullptr)) || ((gt == nullptr) != (gr == nullptr)) || (gt && !tl))
        return -1;
    if (!(prm->radius > 0.0)) return -1;
    islam::SmallArgs a;
    a.N = N; a.E = E; a.M = N - 1; a.B = B;
    int span = 1;
    for (int e = 0; e < E; ++e) {
        const long long i = links_host[2 * e], j = links_host[2 * e + 1];
        if (i < 0 || j < 0 || i >= N || j >= N || i == j) return -2;
        span = std::max<int>(span, (int)std::llabs(i - j));
    }
    a.bw = 9 * (span + 1) - 1; a.span = span;
    a.links = links_dev; a.nodes0 = nodes0; a.vels0 = vels0; a.Z = Z; a.drot = drot; a.dtrans = dtrans; a.dvel = dvel; a.dt = dt;
    for (int k = 0; k < 4; ++k) a.w[k] = w[k];
    a.prm = *prm;
    a.out_nodes = out_nodes; a.out_vels = out_vels; a.out_state = out_state; a.voP = voP; a.tl = tl; a.rl = rl; a.gt = gt; a.gr = gr;
    size_t smem = islam::small_smem_bytes(N, E);
    {
        // speculative retry slots: as many (<= SM_SPEC) as fit behind the window in one SM's shared memory
        static int max_optin = -1;
        if (max_optin < 0) {
            int dev = 0;
            cudaGetDevice(&dev);
            if (cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) max_optin = 0;
        }
        const char* env = std::getenv("ISLAM_SMALL_SPEC");
        const size_t base = (smem + 7) & ~(size_t)7, slot = islam::small_slot_doubles(N, span) * sizeof(double);
        long long fit = span <= 2 ? ((long long)max_optin - 2048 - (long long)base) / (long long)slot : 0;
        if (env) fit = std::min<long long>(fit, std::atoi(env));
        a.nspec = (int)std::max<long long>(0, std::min<long long>(fit, islam::SM_SPEC));
        if (a.nspec > 1) smem = base + a.nspec * slot;
        else a.nspec = 0;
    }
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(islam::k_lm_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    islam::k_lm_small<<<B, islam::SM_THREADS, smem, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

#ifdef ISLAM_PHASE_CLOCKS
extern "C" int islam_debug_small_clocks(long long* out16) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out16, islam::g_small_clk, sizeof(long long) * 16);
}
#endif
