// Kernel family 2 — damped block-sparse Cholesky (multifrontal, level-scheduled) of the LM normal equations.
//
// Replaces the dense `A.diagonal().clamp_; A.diagonal() += A.diagonal()*damping; cholesky_ex; cholesky_solve`
// of PyPose's LM.step + solver.Cholesky as configured at /root/reference/pvgo.py:169-171.
// Same linear system, same clamp and cumulative damping; only the elimination order differs (nested dissection
// over 3-dof variables with trimmed separators, symbolic3.h), so no dense 10N x 10N matrix is ever formed.
//
// One CTA per front.  A front eliminates `npad` pivot variables (Cf = 3 npad columns, a multiple of 9; dummy
// pivots pad the last block) and sees `nb` boundary variables.  Its frontal matrix lives in shared memory as
//   P : (Cf + 3 nb + 1) x Cf column-major panel [F11; F21; rhs^T]   (the right-hand side b = -J^T W r rides along as
//       one extra row, so the forward substitution is part of the factorisation)
//   U : packed lower triangle of the (3 nb + 1)^2 boundary block (update matrix handed to the parent).
// Assembly is "push": the panel is zeroed, the original 3x3 blocks of J^T W J are stored from a per-front list,
// and each child's update matrix is streamed in (coalesced) and added through its boundary -> slot map, one
// child after the other: a fixed summation order, no atomics, bitwise deterministic.
#pragma once
#include "common.cuh"

namespace islam {

#ifdef ISLAM_PHASE_CLOCKS
__device__ long long g_phase_clk[64];
__device__ int g_phase_grid = 1;
#define PHASE(n) do { if (blockIdx.x == 0 && threadIdx.x == 0 && gridDim.x == g_phase_grid) g_phase_clk[n] = clock64(); } while (0)
#else
#define PHASE(n) do { } while (0)
#endif

constexpr int F3_HEAD = 192;     // doubles in front of the panel: 2 x 96 inverse diagonal blocks (also the stage-1 diagonal)

__host__ __device__ __forceinline__ int f3_ld(int Rf) { return (Rf + 3) & ~3; }      // 32-byte aligned columns
__host__ __device__ __forceinline__ long long f3_ulen(int ub) { return (long long)ub * (ub + 1) / 2; }
// packed lower triangle, column-major: element (r, c), r >= c
__host__ __device__ __forceinline__ int f3_uidx(int r, int c, int ub) { return c * ub - c * (c - 1) / 2 + (r - c); }
__host__ __device__ __forceinline__ long long f3_smem_doubles(int Rf, int Cf, int ub, bool u_smem) {
    return F3_HEAD + (long long)f3_ld(Rf) * Cf + 4 + (u_smem ? f3_ulen(ub) : 0);
}

// right-looking 9x9 Cholesky in registers (packed lower A -> L), reciprocal diagonal in linv
__device__ __forceinline__ bool chol9_rl(double* A, double* linv) {
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        double d = A[c * (c + 1) / 2 + c];
        if (!(d > 0.0) || !(d < 1e300)) { ok = false; d = 1.0; }
        double inv = rsqrt(d);
        inv = inv * (1.5 - 0.5 * d * inv * inv);             // one Newton step: full double accuracy
        linv[c] = inv;
        A[c * (c + 1) / 2 + c] = d * inv;
#pragma unroll
        for (int r = c + 1; r < 9; ++r) A[r * (r + 1) / 2 + c] *= inv;
#pragma unroll
        for (int r = c + 1; r < 9; ++r)
#pragma unroll
            for (int c2 = c + 1; c2 <= r; ++c2) A[r * (r + 1) / 2 + c2] -= A[r * (r + 1) / 2 + c] * A[c2 * (c2 + 1) / 2 + c];
    }
    return ok;
}

// column `col` (< 9) of the inverse of the packed lower-triangular L, branch-free: entries above the diagonal
// come out as exact zeros because the partial sums only ever see zeros there
__device__ __forceinline__ void tri_inv_col_uniform(const double* L, const double* linv, int col, double* x) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) s += L[i * (i + 1) / 2 + k] * x[k];
        x[i] = (i == col) ? linv[i] : -s * linv[i];
        if (i < col) x[i] = 0.0;
    }
}

// ---- numeric factorisation of one level ------------------------------------------------------------------------------
// stage 0: local front (original entries + all children) -> factor
// stage 1: multi-GPU, shared front BEFORE the all-reduce: partial frontal matrix (this rank's original entries, undamped,
//          with the partial pivot diagonal kept apart + its private children) dumped to `shared`; no factorisation
// stage 2: multi-GPU, shared front AFTER the all-reduce: `shared` + shared children -> factor
// Per shared front the buffer holds [P compact (Rf x Cf)] [U packed] [original pivot diagonal (Cf)]: PyPose's clamp_ acts
// on the fully summed diagonal of J^T W J before damping (A.4), so the diagonal travels separately.
// Linv: per 9-column block step the inverse of its 9x9 diagonal Cholesky block (row-major), for the back-substitution.
template <int NT, int MINB, bool U_SMEM>
__global__ void __launch_bounds__(NT, MINB)
k_factor3(const LMState* __restrict__ st, const int* __restrict__ fronts, Front3Meta m,
          const double* __restrict__ Hd, const double* __restrict__ Ho, const double* __restrict__ g,
          double* __restrict__ Lbuf, double* __restrict__ Ubuf, double* __restrict__ Linv,
          double* __restrict__ shared, double lm_min_, double lm_max_, double forced_scale, int stage,
          int* chol_fail, const islam_lm_params* __restrict__ prm) {
    // Launched with programmatic stream serialisation (PDL): everything up to cudaGridDependencySynchronize() only
    // touches the immutable symbolic plan and this CTA's shared memory, so it overlaps the tail of the previous level.
    const int f = fronts[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    PHASE(0);
    extern __shared__ double smem[];
    double* sLinv = smem;                      // 2 x 81 (+ pad): double-buffered inverse diagonal blocks
    const int np = m.np[f], npad = m.npad[f], nb = m.nb[f];
    const int Cf = 3 * npad, Rb = 3 * nb, Rf = Cf + Rb + 1, ub = Rb + 1, nbs = npad / 3;
    const int ld = f3_ld(Rf);
    const int ulen = (int)f3_ulen(ub);
    double* P = smem + F3_HEAD;
    double* Lg = Lbuf + m.Loff[f];
    double* Ug = Ubuf + m.Uoff[f];
    double* Uw = U_SMEM ? P + ld * Cf + 4 : Ug;
    const int* vars = m.vars + m.vars_off[f];
    const int k0 = m.child_off[f], nch = m.child_off[f + 1] - k0;
    for (int i = tid; i < ld * Cf + 4; i += NT) P[i] = 0.0;
    if (U_SMEM)
        for (int i = tid; i < ulen; i += NT) Uw[i] = 0.0;
    cudaGridDependencySynchronize();           // previous level (children's U, LM state) complete and visible
    cudaTriggerProgrammaticLaunchCompletion(); // the next level may start its preamble
    if (forced_scale == 0.0 && !st->active) return;
    const double scale = forced_scale != 0.0 ? forced_scale : st->diag_scale;
    const double lm_min = forced_scale != 0.0 ? lm_min_ : prm->lm_min, lm_max = forced_scale != 0.0 ? lm_max_ : prm->lm_max;
    const bool u_accumulates = U_SMEM || nch > 0 || stage != 0;
    if (!U_SMEM && u_accumulates)
        for (int i = tid; i < ulen; i += NT) Ug[i] = 0.0;
    __syncthreads();
    PHASE(1);

    // A1. original entries of J^T W J / -J^T W r first touched by this front (disjoint destinations)
    if (stage != 2) {
        const int o0 = m.orig_off[f], no = m.orig_off[f + 1] - o0;
        for (int idx = tid; idx < 9 * no; idx += NT) {
            const int e = idx / 9, k = idx - 9 * e, c = k / 3, r = k - 3 * c;
            const int rs = m.orig_rs[o0 + e], cs = m.orig_cs[o0 + e], src = m.orig_src[o0 + e];
            if (rs == cs && r < c) continue;                       // diagonal block: lower triangle only
            const double* arr = (src & 2) ? Ho : Hd;
            double v = arr[(size_t)(src >> 2) + ((src & 1) ? 9 * c + r : 9 * r + c)];
            if (rs == cs && r == c) {
                if (stage == 1) { smem[3 * cs + c] = v; v = 0.0; }                         // summed over ranks before the clamp
                else v = fmin(fmax(v, lm_min), lm_max) * scale;                            // clamp, then cumulative damping (A.4)
            }
            P[(3 * rs + r) + (3 * cs + c) * ld] = v;
        }
        for (int idx = tid; idx < 3 * np; idx += NT) P[(Rf - 1) + idx * ld] = -g[3 * (size_t)vars[idx / 3] + idx % 3];
        for (int idx = 3 * np + tid; idx < Cf; idx += NT) {        // dummy pivots: identity, decoupled
            if (stage == 1) smem[idx] = 0.0;
            else P[idx + idx * ld] = 1.0;
        }
    } else {
        const double* base = shared + m.shared_off[f];
        for (int idx = tid; idx < Rf * Cf; idx += NT) {
            const int j = idx / Rf, i = idx - j * Rf;
            double v = base[idx];
            if (i == j) v = (j < 3 * np) ? v + fmin(fmax(base[(size_t)Rf * Cf + ulen + j], lm_min), lm_max) * scale : 1.0;
            P[i + j * ld] = v;
        }
        for (int idx = tid; idx < ulen; idx += NT) Uw[idx] = base[(size_t)Rf * Cf + idx];
    }
    __syncthreads();
    PHASE(2);

    // A2. extend-add of the children's update matrices: one warp per column of the child's packed lower triangle,
    // up to four independent (coalesced) loads in flight per lane
    for (int k = 0; k < nch; ++k) {
        const int c = m.children[k0 + k];
        if (stage == 1 && m.part[c] != m.mypart) continue;         // this rank's private children only
        if (stage == 2 && m.part[c] >= 0) continue;                // shared children only
        const int* cm = m.cmap + m.cmap_off[k0 + k];
        const int ubc = 3 * m.nb[c] + 1;
        const double* Uc = Ubuf + m.Uoff[c];
        for (int cc = warp; cc < ubc - 1; cc += NW) {              // the last column is the unused (rhs, rhs) corner
            const int pc = 3 * cm[cc / 3] + cc % 3;
            const double* col = Uc + ((size_t)cc * ubc - (size_t)cc * (cc - 1) / 2 - cc);
            for (int r0 = cc; r0 < ubc; r0 += 128) {
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int r = r0 + lane + 32 * u; v[u] = r < ubc ? col[r] : 0.0; }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = r0 + lane + 32 * u;
                    if (r >= ubc) continue;
                    const int pr = (r == ubc - 1) ? Rf - 1 : 3 * cm[r / 3] + r % 3;
                    if (pc < Cf) P[pr + pc * ld] += v[u];
                    else Uw[f3_uidx(pr - Cf, pc - Cf, ub)] += v[u];
                }
            }
        }
        __syncthreads();
    }
    PHASE(3);

    if (stage == 1) {                          // dump the partial frontal matrix for the all-reduce
        double* base = shared + m.shared_off[f];
        for (int idx = tid; idx < Rf * Cf; idx += NT) { const int j = idx / Rf, i = idx - j * Rf; base[idx] = P[i + j * ld]; }
        for (int idx = tid; idx < ulen; idx += NT) base[(size_t)Rf * Cf + idx] = Uw[idx];
        for (int idx = tid; idx < Cf; idx += NT) base[(size_t)Rf * Cf + ulen + idx] = smem[idx];
        return;
    }

    // B. right-looking blocked Cholesky of the panel, 9 columns per step, with look-ahead: while warps 1.. apply block
    // column jb to the trailing columns, warp 0 updates just the next 9x9 diagonal block, factors and inverts it (the
    // serial part), so the single-warp latency hides behind the bulk update.  sLinv is double-buffered.
    bool ok = true;
    auto diag_block = [&](int jbn, double* Lout) {           // warp 0 only: P diag block (already updated) -> L, Linv
        const int d0 = 9 * jbn;
        double A[45], linv[9];
#pragma unroll
        for (int r = 0; r < 9; ++r)
#pragma unroll
            for (int q = 0; q <= r; ++q) A[r * (r + 1) / 2 + q] = P[(d0 + r) + (d0 + q) * ld];
        ok = chol9_rl(A, linv) && ok;
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < 9; ++r)
#pragma unroll
                for (int q = 0; q < 9; ++q) P[(d0 + r) + (d0 + q) * ld] = (q <= r) ? A[r * (r + 1) / 2 + (q <= r ? q : 0)] : 0.0;
        }
        if (lane < 9) {
            double x[9];
            tri_inv_col_uniform(A, linv, lane, x);
#pragma unroll
            for (int i = 0; i < 9; ++i) Lout[9 * i + lane] = x[i];
        }
    };
    if (warp == 0) diag_block(0, sLinv);
    __syncthreads();
    for (int jb = 0; jb < nbs; ++jb) {
        const int c0 = 9 * jb;
        const double* sLi = sLinv + 96 * (jb & 1);
        PHASE(10 + 3 * jb);
        if (tid < 81) Linv[m.Ioff[f] + 81 * jb + tid] = sLi[tid];
        // rows below the diagonal block: x = a Lkk^-T
        for (int i = c0 + 9 + tid; i < Rf; i += NT) {
            double acc[9], x[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) acc[q] = P[i + (c0 + q) * ld];
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                double s_ = 0.0;
#pragma unroll
                for (int k = 0; k <= q; ++k) s_ += acc[k] * sLi[9 * q + k];
                x[q] = s_;
            }
#pragma unroll
            for (int q = 0; q < 9; ++q) P[i + (c0 + q) * ld] = x[q];
        }
        __syncthreads();
        PHASE(11 + 3 * jb);
        // trailing update; warp 0 takes the next diagonal block (update + Cholesky + inverse), warps 1.. the rest
        const int ncb = nbs - 1 - jb;
        if (ncb > 0) {
            if (warp == 0) {
                const int d0 = c0 + 9;
                for (int e = lane; e < 45; e += 32) {            // lower triangle of the next diagonal block
                    int r = 0;
                    while ((r + 1) * (r + 2) / 2 <= e) ++r;
                    const int q = e - r * (r + 1) / 2;
                    double s_ = 0.0;
#pragma unroll
                    for (int k = 0; k < 9; ++k) s_ += P[(d0 + r) + (c0 + k) * ld] * P[(d0 + q) + (c0 + k) * ld];
                    P[(d0 + r) + (d0 + q) * ld] -= s_;
                }
                __syncwarp();
                diag_block(jb + 1, sLinv + 96 * ((jb + 1) & 1));
            } else {
                // column block cb only needs rows >= 9 cb (lower trapezoid); the 9 diagonal rows of block jb+1 are warp 0's
                int tasks = 0;
                for (int cb = jb + 1; cb < nbs; ++cb) tasks += 3 * ((Rf - 9 * cb + 3) >> 2);
                for (int t = tid - 32; t < tasks; t += NT - 32) {
                    int cb = jb + 1, rem = t;
                    while (rem >= 3 * ((Rf - 9 * cb + 3) >> 2)) { rem -= 3 * ((Rf - 9 * cb + 3) >> 2); ++cb; }
                    const int S = (Rf - 9 * cb + 3) >> 2;
                    const int c3 = rem / S, rt = rem - c3 * S;
                    const int j0 = 9 * cb, jc = j0 + 3 * c3;
                    int ix[4];
                    bool vx[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        ix[x] = j0 + rt + x * S;
                        vx[x] = ix[x] < Rf && !(cb == jb + 1 && ix[x] < j0 + 9);
                        if (!vx[x]) ix[x] = j0;
                    }
                    double acc[4][3];
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 3; ++y) acc[x][y] = 0.0;
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        const double* col = P + (c0 + q) * ld;
                        double av[4], bv[3];
#pragma unroll
                        for (int x = 0; x < 4; ++x) av[x] = col[ix[x]];
#pragma unroll
                        for (int y = 0; y < 3; ++y) bv[y] = col[jc + y];
#pragma unroll
                        for (int x = 0; x < 4; ++x)
#pragma unroll
                            for (int y = 0; y < 3; ++y) acc[x][y] += av[x] * bv[y];
                    }
#pragma unroll
                    for (int x = 0; x < 4; ++x)
                        if (vx[x])
#pragma unroll
                            for (int y = 0; y < 3; ++y) P[ix[x] + (jc + y) * ld] -= acc[x][y];
                }
            }
        }
        __syncthreads();
        PHASE(12 + 3 * jb);
    }
    if (!ok && tid == 0) *chol_fail = 1;
    PHASE(4);

    // C. keep the factor for the back-substitution (global panel has leading dimension Rf)
    for (int idx = tid; idx < Rf * Cf; idx += NT) {
        const int j = idx / Rf, i = idx - j * Rf;
        Lg[idx] = P[i + j * ld];
    }
#ifdef ISLAM_PHASE_CLOCKS
    __syncthreads();
#endif
    PHASE(5);

    // D. update matrix on the boundary (+ rhs row): U = (children's pass-through) - L21 L21^T.
    // 4x4 register tiles over the lower triangle; operands are LDS.128 pairs: tiles are laid out on ABSOLUTE panel rows
    // from R0 = Cf rounded down to even, so every row quad is 16-byte aligned whatever the parity of Cf (a tile row
    // above Cf is computed and dropped).
    if (ub > 1) {
        const int R0 = Cf & ~1, shift = Cf - R0;
        const int ntr = (ub + shift + 3) >> 2;
        const int ntiles = ntr * (ntr + 1) / 2;
        for (int t = tid; t < ntiles; t += NT) {
            int tr = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
            while ((tr + 1) * (tr + 2) / 2 <= t) ++tr;
            while (tr * (tr + 1) / 2 > t) --tr;
            const int tc = t - tr * (tr + 1) / 2;
            const int r0 = 4 * tr, s0 = 4 * tc;
            double acc[4][4];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
            const double* pa = P + R0 + r0;
            const double* pb = P + R0 + s0;
#pragma unroll 4
            for (int k = 0; k < Cf; ++k) {
                const double2 a01 = *reinterpret_cast<const double2*>(pa + k * ld);
                const double2 a23 = *reinterpret_cast<const double2*>(pa + k * ld + 2);
                const double2 b01 = *reinterpret_cast<const double2*>(pb + k * ld);
                const double2 b23 = *reinterpret_cast<const double2*>(pb + k * ld + 2);
                const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] += av[x] * bv[y];
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const int r = r0 + x - shift, s_ = s0 + y - shift;
                    if (r >= 0 && s_ >= 0 && r < ub && s_ < ub && r >= s_ && !(r == ub - 1 && s_ == ub - 1)) {
                        const int ui = f3_uidx(r, s_, ub);
                        if (u_accumulates) Uw[ui] -= acc[x][y];
                        else Uw[ui] = -acc[x][y];
                    }
                }
        }
        if (U_SMEM) {
            __syncthreads();
            for (int i = tid; i < ulen; i += NT) Ug[i] = Uw[i];
        }
    }
#ifdef ISLAM_PHASE_CLOCKS
    __syncthreads();
#endif
    PHASE(6);
}

// ---- back-substitution of one level (root first) ---------------------------------------------------------------------
// x_p = L11^-T (y_p - L21^T x_b), blocked by 9 columns: the inverse 9x9 diagonal blocks were stored by the factor
// kernel, so every block step is a 9x9 mat-vec followed by a 9-deep update of the earlier unknowns.
// The whole panel is pulled into shared memory with one burst of independent loads first: the factor was written a
// few hundred MB of traffic ago, so every access is a DRAM-latency access and must not sit on a dependent chain.
constexpr int BS3_THREADS = 512;
__host__ __device__ __forceinline__ long long bs3_smem_doubles(int Rf, int Cf, bool staged) {
    const int Rb = Rf - Cf - 1;
    return Rb + Cf + 16 + 9LL * Cf + (staged ? (long long)Rf * Cf : 0);
}

__global__ void __launch_bounds__(BS3_THREADS)
k_backsolve3(const LMState* __restrict__ st, const int* __restrict__ fronts, Front3Meta m,
             const double* __restrict__ Lbuf, const double* __restrict__ Linv, double* __restrict__ D, int force,
             int smem_doubles) {
    const int f = fronts[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    constexpr int NW = BS3_THREADS / 32;
    extern __shared__ double smem[];
    const int np = m.np[f], npad = m.npad[f], nb = m.nb[f];
    const int Cf = 3 * npad, Rb = 3 * nb, Rf = Cf + Rb + 1, ld = Rf, nbs = npad / 3;
    const int* vars = m.vars + m.vars_off[f];
    const double* Lg = Lbuf + m.Loff[f];
    double* xb = smem;                 // [Rb]
    double* ts = xb + Rb;              // [Cf]
    double* xs = ts + Cf;              // [16]
    double* sLi = xs + 16;             // [nbs][81] inverse diagonal blocks
    double* sP = sLi + 81 * nbs;       // [Rf x Cf] panel copy (if it fits)
    const bool staged = (bs3_smem_doubles(Rf, Cf, true) <= smem_doubles);
    cudaGridDependencySynchronize();           // PDL: the parents' solution (and, for the root, the factor) is complete
    cudaTriggerProgrammaticLaunchCompletion();
    if (!force && !st->active) return;
    for (int r = tid; r < Rb; r += BS3_THREADS) xb[r] = D[3 * (size_t)vars[npad + r / 3] + (r % 3)];
    for (int i = tid; i < 81 * nbs; i += BS3_THREADS) sLi[i] = Linv[m.Ioff[f] + i];
    if (staged)
        for (int i = tid; i < Rf * Cf; i += BS3_THREADS) sP[i] = Lg[i];
    const double* Lp = staged ? sP : Lg;
    __syncthreads();
    // ts[c] = y[c] - sum_r L21[r,c] xb[r]      (one warp per column)
    for (int c = w; c < Cf; c += NW) {
        const double* col = Lp + (size_t)c * ld + Cf;
        double s = 0.0;
        for (int r = lane; r < Rb; r += 32) s += col[r] * xb[r];
        s = warp_sum(s);
        if (lane == 0) ts[c] = col[Rb] - s;       // rhs row holds y = L11^-1 (b - ...)
    }
    __syncthreads();
    for (int jb = nbs - 1; jb >= 0; --jb) {
        const int c0 = 9 * jb;
        if (tid < 9) {                             // x_blk = Lkk^-T ts_blk :  x[a] = sum_{b>=a} Linv[b][a] ts[b]
            const double* Li = sLi + 81 * jb;
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < 9; ++b) s += (b >= tid) ? Li[9 * b + tid] * ts[c0 + b] : 0.0;
            xs[tid] = s;
        }
        __syncthreads();
        if (tid < 9) ts[c0 + tid] = xs[tid];
        for (int k = tid; k < c0; k += BS3_THREADS) {   // ts[k] -= sum_{c in blk} L[c,k] x[c]   (column k, rows c0..c0+8)
            const double* col = Lp + (size_t)k * ld + c0;
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < 9; ++q) s += col[q] * xs[q];
            ts[k] -= s;
        }
        __syncthreads();
    }
    for (int c = tid; c < 3 * np; c += BS3_THREADS) D[3 * (size_t)vars[c / 3] + (c % 3)] = ts[c];
}

}  // namespace islam
