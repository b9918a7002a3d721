// Kernel family 2 — damped block-sparse Cholesky (multifrontal, level-scheduled) of the LM normal equations.
//
// Replaces the dense `A.diagonal().clamp_; A.diagonal() += A.diagonal()*damping; cholesky_ex; cholesky_solve`
// of PyPose's LM.step + solver.Cholesky as configured at /root/reference/pvgo.py:169-171.
// Same linear system, same clamp and cumulative damping; only the elimination order differs (nested
// dissection over the pose index, symbolic.h), so no dense 10N x 10N matrix is ever formed.
//
// One CTA per front.  A front owns `np` pivot poses (9 np columns) and sees `nb` boundary poses; its panel is
// the (9(np+nb)+1) x 9np lower trapezoid [F11; F21; rhs^T], column-major, staged in shared memory.  The right-
// hand side b = -J^T W r rides along as one extra row, so the forward substitution is part of the
// factorisation.  All sums are "pull" gathers in a fixed order: bitwise deterministic, no atomics.
#pragma once
#include "common.cuh"

namespace islam {

constexpr int FAC_THREADS = 256;
constexpr int BS_THREADS = 128;

// ---- gather of the children's update matrices ------------------------------------------------------------
// rs: parent row slot (or -1 for the rhs row), a: dof in slot; cs/b likewise for the column
__device__ __forceinline__ double pull_children(const FrontMeta& m, const double* __restrict__ Ubuf, int f, int ns,
                                                int rs, int a, int cs, int b, int mode) {
    double v = 0.0;
    int k0 = m.child_off[f], k1 = m.child_off[f + 1];
    for (int k = k0; k < k1; ++k) {
        int c = m.children[k];
        if (mode == 1 && m.part[c] >= 0) continue;     // shared children only
        if (mode == 2 && m.part[c] < 0) continue;      // private children only
        const int* inv = m.cinv + m.cinv_off[k];
        int nbc = m.nb[c];
        int cc = inv[cs];
        if (cc < 0) continue;
        int rc;
        if (rs < 0) rc = 9 * nbc;
        else {
            int t = inv[rs];
            if (t < 0) continue;
            rc = 9 * t + a;
        }
        int ldu = 9 * nbc + 1;
        v += Ubuf[m.Uoff[c] + rc + (long long)(9 * cc + b) * ldu];
    }
    return v;
}

// original (undamped) H / b entries of a panel element; clamp + cumulative damping on pivot diagonals (A.4)
__device__ __forceinline__ double orig_entry(const FrontMeta& m, int f, int np, const int* __restrict__ nodes,
                                             int rs, int a, int cs, int b, const double* __restrict__ Hd,
                                             const double* __restrict__ Ho, const double* __restrict__ g,
                                             double scale, double lm_min, double lm_max, bool damp) {
    int nc = nodes[cs];
    if (rs < 0) return -g[9 * (size_t)nc + b];                      // b = -J^T W r
    if (rs == cs) {
        double v = Hd[81 * (size_t)nc + 9 * a + b];
        if (a == b && damp) v = fmin(fmax(v, lm_min), lm_max) * scale;
        return v;
    }
    int h = m.hmap[m.hmap_off[f] + rs * np + cs];
    if (h < 0) return 0.0;
    const double* blk = Ho + 81 * (size_t)(h >> 1);
    return (h & 1) ? blk[9 * b + a] : blk[9 * a + b];
}

// 9x9 Cholesky of the diagonal block, fully unrolled in registers; every thread does it redundantly
__device__ __forceinline__ bool chol9(const double* __restrict__ Dblk, double* Lk, double* linv) {
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 9; ++c) {
#pragma unroll
        for (int k = 0; k <= c; ++k) {
            double s = Dblk[9 * c + k];
#pragma unroll
            for (int q = 0; q < k; ++q) s -= Lk[c * (c + 1) / 2 + q] * Lk[k * (k + 1) / 2 + q];
            if (k == c) {
                if (!(s > 0.0) || !(s < 1e300)) { ok = false; s = 1.0; }
                double d = sqrt(s);
                Lk[c * (c + 1) / 2 + c] = d;
                linv[c] = 1.0 / d;
            } else {
                Lk[c * (c + 1) / 2 + k] = s * linv[k];
            }
        }
    }
    return ok;
}

// ---- numeric factorisation of one level ------------------------------------------------------------------
// stage: 0 = all fronts of this launch are local (assemble from H and all children)
//        1 = "base" pass of shared fronts (multi-GPU): write orig + private-children sums into `shared`
//        2 = shared fronts after the all-reduce: assemble from `shared` + shared children
__global__ void __launch_bounds__(FAC_THREADS, 1)
k_factor_level(const LMState* __restrict__ st, const int* __restrict__ fronts, FrontMeta m,
               const double* __restrict__ Hd, const double* __restrict__ Ho, const double* __restrict__ g,
               double* __restrict__ Lbuf, double* __restrict__ Ubuf, const double* __restrict__ shared,
               double lm_min, double lm_max, double forced_scale, int smem_doubles, int stage, int* chol_fail) {
    if (forced_scale == 0.0 && !st->active) return;
    const double scale = forced_scale != 0.0 ? forced_scale : st->diag_scale;
    const int f = fronts[blockIdx.x];
    const int tid = threadIdx.x;
    extern __shared__ double smem[];
    const int np = m.np[f], nb = m.nb[f], ns = np + nb;
    const int Cf = 9 * np, Rb = 9 * nb, Rf = Cf + Rb + 1, ld = Rf, ub = Rb + 1;
    const int* nodes = m.nodes + m.nodes_off[f];
    double* Lg = Lbuf + m.Loff[f];
    double* Ug = Ubuf + m.Uoff[f];
    const bool in_smem = (Rf * Cf + 96 <= smem_doubles);
    double* Dblk = smem;                       // 81 doubles
    double* P = in_smem ? smem + 96 : Lg;
    const double* base = (stage == 2) ? shared + m.shared_off[f] : nullptr;   // Rf x Rf column-major

    // A. assemble the panel
    for (int idx = tid; idx < Rf * Cf; idx += FAC_THREADS) {
        int j = idx / Rf, i = idx - j * Rf;
        double v = 0.0;
        if (i >= j) {
            int cs = j / 9, b = j - 9 * cs;
            int rs = (i == Rf - 1) ? -1 : i / 9;
            int a = (rs < 0) ? 0 : i - 9 * rs;
            if (stage == 2) {
                v = base[i + (long long)j * Rf];
                if (i == j)                     // all-reduced original diagonal: clamp, then cumulative damping (A.4)
                    v += fmin(fmax(base[(long long)Rf * Rf + j], lm_min), lm_max) * scale;
                v += pull_children(m, Ubuf, f, ns, rs, a, cs, b, 1);
            } else {
                v = orig_entry(m, f, np, nodes, rs, a, cs, b, Hd, Ho, g, scale, lm_min, lm_max, true);
                v += pull_children(m, Ubuf, f, ns, rs, a, cs, b, 0);
            }
        }
        P[i + (long long)j * ld] = v;
    }
    __syncthreads();

    // B. left-looking blocked Cholesky, 9 columns (one pose) at a time
    bool ok = true;
    for (int jb = 0; jb < np; ++jb) {
        const int c0 = 9 * jb;
        // (a) apply the already-final columns 0..c0-1 to block column jb
        for (int i = c0 + tid; i < Rf; i += FAC_THREADS) {
            double acc[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) acc[c] = P[i + (long long)(c0 + c) * ld];
            for (int k = 0; k < c0; ++k) {
                double a = P[i + (long long)k * ld];
                const double* col = P + (long long)k * ld + c0;
#pragma unroll
                for (int c = 0; c < 9; ++c) acc[c] -= a * col[c];
            }
            if (i < c0 + 9) {
#pragma unroll
                for (int c = 0; c < 9; ++c) Dblk[9 * (i - c0) + c] = acc[c];
            } else {
#pragma unroll
                for (int c = 0; c < 9; ++c) P[i + (long long)(c0 + c) * ld] = acc[c];
            }
        }
        __syncthreads();
        // (b) factor the 9x9 diagonal block (redundantly per thread) and solve the rows below it
        double Lk[45], linv[9];
        ok = chol9(Dblk, Lk, linv) && ok;
        for (int i = c0 + tid; i < Rf; i += FAC_THREADS) {
            if (i < c0 + 9) {
                int rr = i - c0;
#pragma unroll
                for (int c = 0; c < 9; ++c)
#pragma unroll
                    for (int r2 = 0; r2 < 9; ++r2)
                        if (r2 == rr) P[i + (long long)(c0 + c) * ld] = (c <= r2) ? Lk[r2 * (r2 + 1) / 2 + (c <= r2 ? c : 0)] : 0.0;
            } else {
                double x[9];
#pragma unroll
                for (int c = 0; c < 9; ++c) {
                    double s = P[i + (long long)(c0 + c) * ld];
#pragma unroll
                    for (int k = 0; k < c; ++k) s -= x[k] * Lk[c * (c + 1) / 2 + k];
                    x[c] = s * linv[c];
                }
#pragma unroll
                for (int c = 0; c < 9; ++c) P[i + (long long)(c0 + c) * ld] = x[c];
            }
        }
        __syncthreads();
    }
    if (!ok && tid == 0) *chol_fail = 1;

    // C. keep the factor for the back-substitution
    if (in_smem)
        for (int idx = tid; idx < Rf * Cf; idx += FAC_THREADS) Lg[idx] = P[idx];

    // D. update matrix  U = (children's pass-through) - L21 L21^T  on the boundary (+ rhs row), lower triangle
    if (ub > 1) {
        const double* L21 = P + Cf;
        const int ntr = (ub + 3) >> 2;
        const int ntiles = ntr * (ntr + 1) / 2;
        for (int t = tid; t < ntiles; t += FAC_THREADS) {
            int tr = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
            while ((tr + 1) * (tr + 2) / 2 <= t) ++tr;
            while (tr * (tr + 1) / 2 > t) --tr;
            int tc = t - tr * (tr + 1) / 2;
            int r0 = 4 * tr, s0 = 4 * tc;
            int ri[4], si[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { ri[q] = min(r0 + q, ub - 1); si[q] = min(s0 + q, ub - 1); }
            double acc[4][4];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
            for (int k = 0; k < Cf; ++k) {
                const double* col = L21 + (long long)k * ld;
                double av[4], bv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { av[q] = col[ri[q]]; bv[q] = col[si[q]]; }
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] += av[x] * bv[y];
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    int r = r0 + x, s = s0 + y;
                    if (r < ub && s < ub && r >= s && !(r == ub - 1 && s == ub - 1)) {
                        int cs = np + s / 9, b = s % 9;
                        int rs = (r == ub - 1) ? -1 : np + r / 9;
                        int a = (rs < 0) ? 0 : r % 9;
                        double v;
                        if (stage == 2)
                            v = base[(Cf + r) + (long long)(Cf + s) * Rf] + pull_children(m, Ubuf, f, ns, rs, a, cs, b, 1);
                        else
                            v = pull_children(m, Ubuf, f, ns, rs, a, cs, b, 0);
                        Ug[r + (long long)s * ub] = v - acc[x][y];
                    }
                }
        }
    }
}

// ---- multi-GPU: pre-all-reduce base of the shared fronts ------------------------------------------------
// Per shared front: Rf x Rf column-major partial sums (original entries of the factors this rank owns + update
// matrices of its private children) followed by the 9np partial ORIGINAL pivot diagonals, kept apart because
// PyPose's clamp_ acts on the fully summed diagonal of J^T W J before damping (A.4).
__global__ void __launch_bounds__(FAC_THREADS)
k_shared_base(const LMState* __restrict__ st, const int* __restrict__ fronts, FrontMeta m,
              const double* __restrict__ Hd, const double* __restrict__ Ho, const double* __restrict__ g,
              const double* __restrict__ Ubuf, double* __restrict__ shared) {
    if (!st->active) return;
    const int f = fronts[blockIdx.x];
    const int np = m.np[f], nb = m.nb[f], ns = np + nb;
    const int Cf = 9 * np, Rf = 9 * ns + 1;
    const int* nodes = m.nodes + m.nodes_off[f];
    double* base = shared + m.shared_off[f];
    for (long long idx = threadIdx.x; idx < (long long)Rf * Rf; idx += FAC_THREADS) {
        int j = (int)(idx / Rf), i = (int)(idx - (long long)j * Rf);
        double v = 0.0;
        if (i >= j && j < Rf - 1) {
            int cs = j / 9, b = j - 9 * cs;
            int rs = (i == Rf - 1) ? -1 : i / 9;
            int a = (rs < 0) ? 0 : i - 9 * rs;
            if (j < Cf) {
                double o = orig_entry(m, f, np, nodes, rs, a, cs, b, Hd, Ho, g, 1.0, 0.0, 0.0, false);
                if (i == j) base[(long long)Rf * Rf + j] = o;
                else v = o;
            }
            v += pull_children(m, Ubuf, f, ns, rs, a, cs, b, 2);
        }
        base[idx] = v;
    }
}

// ---- back-substitution of one level (root first) --------------------------------------------------------
__global__ void __launch_bounds__(BS_THREADS)
k_backsolve_level(const LMState* __restrict__ st, const int* __restrict__ fronts, FrontMeta m,
                  const double* __restrict__ Lbuf, double* __restrict__ D, int force) {
    if (!force && !st->active) return;
    const int f = fronts[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    extern __shared__ double smem[];
    const int np = m.np[f], nb = m.nb[f];
    const int Cf = 9 * np, Rb = 9 * nb, Rf = Cf + Rb + 1, ld = Rf, ldt = Cf + 1;
    const int* nodes = m.nodes + m.nodes_off[f];
    const double* Lg = Lbuf + m.Loff[f];
    double* xb = smem;                 // [Rb]
    double* ts = xb + Rb;              // [Cf]
    double* LT = ts + Cf;              // [Cf x ldt] L11 transposed: LT[k + c*ldt] = L[c,k]
    for (int r = tid; r < Rb; r += BS_THREADS) {
        int slot = np + r / 9;
        xb[r] = D[9 * (size_t)nodes[slot] + (r % 9)];
    }
    for (int idx = tid; idx < Cf * Cf; idx += BS_THREADS) {
        int k = idx / Cf, c = idx - k * Cf;      // element L[c,k], c fastest (coalesced global read)
        LT[k + c * ldt] = (c >= k) ? Lg[c + (long long)k * ld] : 0.0;
    }
    __syncthreads();
    // ts[c] = y[c] - sum_r L21[r,c] xb[r]
    for (int c = w; c < Cf; c += BS_THREADS / 32) {
        const double* col = Lg + (long long)c * ld + Cf;
        double s = 0.0;
        for (int r = lane; r < Rb; r += 32) s += col[r] * xb[r];
        s = warp_sum(s);
        if (lane == 0) ts[c] = col[Rb] - s;       // rhs row holds y = L11^-1 (b - ...)
    }
    __syncthreads();
    // L11^T x = ts, column-oriented, one warp
    if (w == 0) {
        for (int c = Cf - 1; c >= 0; --c) {
            double xc = ts[c] / LT[c + c * ldt];
            __syncwarp();
            if (lane == 0) ts[c] = xc;
            for (int k = lane; k < c; k += 32) ts[k] -= LT[k + c * ldt] * xc;
            __syncwarp();
        }
    }
    __syncthreads();
    for (int c = tid; c < Cf; c += BS_THREADS) D[9 * (size_t)nodes[c / 9] + (c % 9)] = ts[c];
}

}  // namespace islam
