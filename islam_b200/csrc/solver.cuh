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
        if (mode == 2 && m.part[c] != m.mypart) continue;   // this rank's private children only
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

// ---- per-front metadata staged in shared memory -------------------------------------------------------------
// The gathers below touch the maps once per matrix element; chasing them through L2 (5-6 dependent loads per
// element) made a single front cost ~240 us, so a CTA copies everything it needs into shared memory first.
constexpr int MAXC = 4;      // children handled on the fast path (more => generic global-memory path)
constexpr int MAXS = 40;     // slots (pivot + boundary poses) handled on the fast path

struct FrontCtx {
    int np, nb, ns, nch;
    bool fast;
    const int* nodes;        // [ns]            (shared or global)
    const int* hmap;         // [ns*np]
    const int* inv;          // [nch][ns]       parent slot -> child boundary idx | -1   (fast path only)
    const int* cnb;          // [nch]
    const int* cshared;      // [nch] 1 if the child is a shared (multi-GPU) front
    const long long* cU;     // [nch] offset of the child's U
};

__host__ __device__ __forceinline__ int front_meta_doubles(int np, int ns, int nch) {
    if (nch > MAXC || ns > MAXS) return 0;
    int ints = 2 * MAXC /* long long offsets */ + ns + ns * np + nch * ns + 2 * nch;
    return (((ints + 1) / 2 + 1) + 1) & ~1;       // even: the panel behind it stays 16-byte aligned
}

__device__ __forceinline__ void stage_front(const FrontMeta& m, int f, int* si, FrontCtx& c) {
    c.np = m.np[f]; c.nb = m.nb[f]; c.ns = c.np + c.nb;
    const int k0 = m.child_off[f];
    c.nch = m.child_off[f + 1] - k0;
    c.fast = (c.nch <= MAXC && c.ns <= MAXS);
    const int* gn = m.nodes + m.nodes_off[f];
    const int* gh = m.hmap + m.hmap_off[f];
    if (!c.fast) { c.nodes = gn; c.hmap = gh; c.inv = nullptr; c.cnb = nullptr; c.cshared = nullptr; c.cU = nullptr; return; }
    long long* sU = (long long*)si;                 // 8-byte aligned: first in the region
    int* sn = si + 2 * MAXC;
    int* sh = sn + c.ns;
    int* sv = sh + c.ns * c.np;
    int* sb = sv + c.nch * c.ns;
    int* ss = sb + c.nch;
    for (int i = threadIdx.x; i < c.ns; i += blockDim.x) sn[i] = gn[i];
    for (int i = threadIdx.x; i < c.ns * c.np; i += blockDim.x) sh[i] = gh[i];
    for (int i = threadIdx.x; i < c.nch * c.ns; i += blockDim.x) {
        int k = i / c.ns, s_ = i - k * c.ns;
        sv[i] = m.cinv[m.cinv_off[k0 + k] + s_];
    }
    for (int k = threadIdx.x; k < c.nch; k += blockDim.x) {
        int ch = m.children[k0 + k];
        sb[k] = m.nb[ch];
        ss[k] = m.part[ch] < 0;
        sU[k] = m.Uoff[ch];
    }
    c.nodes = sn; c.hmap = sh; c.inv = sv; c.cnb = sb; c.cshared = ss; c.cU = sU;
}

// gather of the children's update matrices at (row slot rs | -1 = rhs, dof a ; col slot cs, dof b)
// mode 0: all children, 1: shared children only, 2: private children only
__device__ __forceinline__ double pull(const FrontCtx& c, const FrontMeta& m, const double* __restrict__ Ubuf, int f,
                                       int rs, int a, int cs, int b, int mode) {
    if (!c.fast) return pull_children(m, Ubuf, f, c.ns, rs, a, cs, b, mode);
    double v = 0.0;
#pragma unroll 2
    for (int k = 0; k < c.nch; ++k) {
        if (mode == 1 && !c.cshared[k]) continue;
        if (mode == 2 && c.cshared[k]) continue;
        const int* inv = c.inv + k * c.ns;
        int cc = inv[cs];
        if (cc < 0) continue;
        int nbc = c.cnb[k];
        int rc;
        if (rs < 0) rc = 9 * nbc;
        else {
            int t = inv[rs];
            if (t < 0) continue;
            rc = 9 * t + a;
        }
        v += Ubuf[c.cU[k] + rc + (long long)(9 * cc + b) * (9 * nbc + 1)];
    }
    return v;
}

__device__ __forceinline__ double orig(const FrontCtx& c, int rs, int a, int cs, int b, const double* __restrict__ Hd,
                                       const double* __restrict__ Ho, const double* __restrict__ g, double scale,
                                       double lm_min, double lm_max, bool damp) {
    int nc = c.nodes[cs];
    if (rs < 0) return -g[9 * (size_t)nc + b];                      // b = -J^T W r
    if (rs == cs) {
        double v = Hd[81 * (size_t)nc + 9 * a + b];
        if (a == b && damp) v = fmin(fmax(v, lm_min), lm_max) * scale;
        return v;
    }
    int h = c.hmap[rs * c.np + cs];
    if (h < 0) return 0.0;
    const double* blk = Ho + 81 * (size_t)(h >> 1);
    return (h & 1) ? blk[9 * b + a] : blk[9 * a + b];
}

// column `col` of the inverse of the 9x9 lower-triangular Lk (forward substitution of e_col), unrolled
__device__ __forceinline__ void tri_inv_col(const double* Lk, const double* linv, int col, double* __restrict__ out /*[9][9] row-major*/) {
#pragma unroll
    for (int c0 = 0; c0 < 9; ++c0) {
        if (c0 == col) {
            double x[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                if (i < c0) { x[i] = 0.0; continue; }
                if (i == c0) { x[i] = linv[i]; continue; }
                double s = 0.0;
#pragma unroll
                for (int k = c0; k < i; ++k) s += Lk[i * (i + 1) / 2 + k] * x[k];
                x[i] = -s * linv[i];
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) out[9 * i + c0] = x[i];
        }
    }
}

// ---- numeric factorisation of one level ------------------------------------------------------------------
// stage: 0 = local fronts (assemble from H and all children)
//        2 = shared fronts after the all-reduce (multi-GPU): assemble from `shared` + shared children
// Linv: per pivot pose the inverse of its 9x9 diagonal Cholesky block (row-major), for the back-substitution.
__global__ void __launch_bounds__(FAC_THREADS, 1)
k_factor_level(const LMState* __restrict__ st, const int* __restrict__ fronts, FrontMeta m,
               const double* __restrict__ Hd, const double* __restrict__ Ho, const double* __restrict__ g,
               double* __restrict__ Lbuf, double* __restrict__ Ubuf, double* __restrict__ Linv,
               const double* __restrict__ shared, double lm_min_, double lm_max_, double forced_scale, int smem_doubles,
               int stage, int* chol_fail, const islam_lm_params* __restrict__ prm) {
    if (forced_scale == 0.0 && !st->active) return;
    const double scale = forced_scale != 0.0 ? forced_scale : st->diag_scale;
    const double lm_min = forced_scale != 0.0 ? lm_min_ : prm->lm_min, lm_max = forced_scale != 0.0 ? lm_max_ : prm->lm_max;
    const int f = fronts[blockIdx.x];
    const int tid = threadIdx.x;
    extern __shared__ double smem[];
    double* Dblk = smem;                       // 81 doubles (+ pad)
    FrontCtx c;
    stage_front(m, f, (int*)(smem + 96), c);
    const int np = c.np, nb = c.nb;
    const int Cf = 9 * np, Rb = 9 * nb, Rf = Cf + Rb + 1, ld = Rf, ub = Rb + 1;
    const int meta_doubles = front_meta_doubles(np, c.ns, c.nch);
    double* Lg = Lbuf + m.Loff[f];
    double* Ug = Ubuf + m.Uoff[f];
    const bool in_smem = (96 + meta_doubles + Rf * Cf <= smem_doubles);
    double* P = in_smem ? smem + 96 + meta_doubles : Lg;
    const double* base = (stage == 2) ? shared + m.shared_off[f] : nullptr;   // Rf x Rf column-major (+ Cf diag)
    __syncthreads();

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
                v += pull(c, m, Ubuf, f, rs, a, cs, b, 1);
            } else {
                v = orig(c, rs, a, cs, b, Hd, Ho, g, scale, lm_min, lm_max, true);
                v += pull(c, m, Ubuf, f, rs, a, cs, b, 0);
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
            for (int q = 0; q < 9; ++q) acc[q] = P[i + (long long)(c0 + q) * ld];
            for (int k = 0; k < c0; ++k) {
                double a = P[i + (long long)k * ld];
                const double* col = P + (long long)k * ld + c0;
#pragma unroll
                for (int q = 0; q < 9; ++q) acc[q] -= a * col[q];
            }
            if (i < c0 + 9) {
#pragma unroll
                for (int q = 0; q < 9; ++q) Dblk[9 * (i - c0) + q] = acc[q];
            } else {
#pragma unroll
                for (int q = 0; q < 9; ++q) P[i + (long long)(c0 + q) * ld] = acc[q];
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
                for (int q = 0; q < 9; ++q)
#pragma unroll
                    for (int r2 = 0; r2 < 9; ++r2)
                        if (r2 == rr) P[i + (long long)(c0 + q) * ld] = (q <= r2) ? Lk[r2 * (r2 + 1) / 2 + (q <= r2 ? q : 0)] : 0.0;
                tri_inv_col(Lk, linv, rr, Linv + 81 * (size_t)c.nodes[jb]);
            } else {
                double x[9];
#pragma unroll
                for (int q = 0; q < 9; ++q) {
                    double s_ = P[i + (long long)(c0 + q) * ld];
#pragma unroll
                    for (int k = 0; k < q; ++k) s_ -= x[k] * Lk[q * (q + 1) / 2 + k];
                    x[q] = s_ * linv[q];
                }
#pragma unroll
                for (int q = 0; q < 9; ++q) P[i + (long long)(c0 + q) * ld] = x[q];
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
                    int r = r0 + x, s_ = s0 + y;
                    if (r < ub && s_ < ub && r >= s_ && !(r == ub - 1 && s_ == ub - 1)) {
                        int cs = np + s_ / 9, b = s_ % 9;
                        int rs = (r == ub - 1) ? -1 : np + r / 9;
                        int a = (rs < 0) ? 0 : r % 9;
                        double v;
                        if (stage == 2)
                            v = base[(Cf + r) + (long long)(Cf + s_) * Rf] + pull(c, m, Ubuf, f, rs, a, cs, b, 1);
                        else
                            v = pull(c, m, Ubuf, f, rs, a, cs, b, 0);
                        Ug[r + (long long)s_ * ub] = v - acc[x][y];
                    }
                }
        }
    }
}

// ---- fast path: one CTA of 512 threads per front, everything block-structured ------------------------------------
// Used when every front of a level fits the shared-memory budget with <= MAXC children and <= MAXS slots (always
// the case for chain / band graphs; fronts that do not qualify take k_factor_level above).
constexpr int F3_THREADS = 512;
#ifdef ISLAM_PHASE_CLOCKS
__device__ long long g_phase_clk[64];
__device__ int g_phase_grid = 1;
#define PHASE(n) do { if (blockIdx.x == 0 && threadIdx.x == 0 && gridDim.x == g_phase_grid) g_phase_clk[n] = clock64(); } while (0)
#else
#define PHASE(n) do { } while (0)
#endif

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

// column `lane` (< 9) of the inverse of the packed lower-triangular L, branch-free: entries above the diagonal
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

__global__ void __launch_bounds__(F3_THREADS, 1)
k_factor_fast(const LMState* __restrict__ st, const int* __restrict__ fronts, FrontMeta m,
              const double* __restrict__ Hd, const double* __restrict__ Ho, const double* __restrict__ g,
              double* __restrict__ Lbuf, double* __restrict__ Ubuf, double* __restrict__ Linv,
              const double* __restrict__ shared, double lm_min_, double lm_max_, double forced_scale, int stage,
              int* chol_fail, const islam_lm_params* __restrict__ prm) {
    // Launched with programmatic stream serialisation (PDL): everything up to cudaGridDependencySynchronize() only
    // touches the immutable symbolic plan, so it overlaps the tail of the previous level's kernel.
    const int f = fronts[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = F3_THREADS / 32;
    PHASE(0);
    extern __shared__ double smem[];
    double* sLinv = smem;                      // 2 x 81 (+ pad): double-buffered inverse diagonal blocks
    FrontCtx c;
    stage_front(m, f, (int*)(smem + 192), c);
    const int np = c.np, nb = c.nb, ns = c.ns;
    const int Cf = 9 * np, Rb = 9 * nb, Rf = Cf + Rb + 1, ub = Rb + 1;
    const int ld = (Rf + 3) & ~3;              // 32-byte aligned columns: LDS.128 on row quads
    double* P = smem + 192 + front_meta_doubles(np, ns, c.nch);
    double* Lg = Lbuf + m.Loff[f];
    double* Ug = Ubuf + m.Uoff[f];
    const double* base = (stage == 2) ? shared + m.shared_off[f] : nullptr;
    cudaGridDependencySynchronize();           // previous level (children's U, LM state) complete and visible
    cudaTriggerProgrammaticLaunchCompletion(); // the next level may start staging its metadata
    if (forced_scale == 0.0 && !st->active) return;
    const double scale = forced_scale != 0.0 ? forced_scale : st->diag_scale;
    const double lm_min = forced_scale != 0.0 ? lm_min_ : prm->lm_min, lm_max = forced_scale != 0.0 ? lm_max_ : prm->lm_max;
    __syncthreads();
    PHASE(1);

    // A. assembly, one warp per 9x9 block (row block rs in [cs, ns]; rs == ns is the rhs row); two blocks are in
    // flight per warp so that their (DRAM-latency) gathers overlap
    {
        int ea[3], eb[3];                      // this lane's (row, col) inside a 9x9 block, row fastest
#pragma unroll
        for (int j = 0; j < 3; ++j) { int e = lane + 32 * j; eb[j] = e / 9; ea[j] = e - 9 * eb[j]; }
        int nblk = 0;
        for (int cs = 0; cs < np; ++cs) nblk += ns + 1 - cs;
        for (int blk0 = warp; blk0 < nblk; blk0 += 2 * NW) {
            if (blk0 == warp) PHASE(40); else if (blk0 == warp + 2 * NW) PHASE(43);
            double v[2][3], tv[2][MAXC][3];
            int i0s[2], j0s[2];
            bool rhss[2], valid[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int blk = blk0 + u * NW;
                valid[u] = blk < nblk;
                int cs = 0, rem = valid[u] ? blk : 0;
                while (rem >= ns + 1 - cs) { rem -= ns + 1 - cs; ++cs; }
                const int rs = cs + rem;
                const bool rhs = (rs == ns);
                const int nc = c.nodes[cs];
                const int i0 = rhs ? Rf - 1 : 9 * rs, j0 = 9 * cs;
                i0s[u] = i0; j0s[u] = j0; rhss[u] = rhs;
#pragma unroll
                for (int j = 0; j < 3; ++j) v[u][j] = 0.0;
                if (stage == 2) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        int e = lane + 32 * j;
                        if (e < (rhs ? 9 : 81)) {
                            int a = rhs ? 0 : ea[j], b = rhs ? e : eb[j];
                            v[u][j] = base[(i0 + a) + (long long)(j0 + b) * Rf];
                            if (i0 + a == j0 + b) v[u][j] += fmin(fmax(base[(long long)Rf * Rf + j0 + b], lm_min), lm_max) * scale;
                        }
                    }
                } else if (rhs) {
                    if (lane < 9) v[u][0] = -g[9 * (size_t)nc + lane];
                } else if (rs == cs) {
                    const double* ob = Hd + 81 * (size_t)nc;
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        if (lane + 32 * j < 81) {
                            double t = ob[9 * ea[j] + eb[j]];
                            v[u][j] = (ea[j] == eb[j]) ? fmin(fmax(t, lm_min), lm_max) * scale : t;
                        }
                } else {
                    int h = c.hmap[rs * np + cs];
                    if (h >= 0) {
                        const double* ob = Ho + 81 * (size_t)(h >> 1);
                        const bool tr = h & 1;
#pragma unroll
                        for (int j = 0; j < 3; ++j)
                            if (lane + 32 * j < 81) v[u][j] = tr ? ob[9 * eb[j] + ea[j]] : ob[9 * ea[j] + eb[j]];
                    }
                }
#pragma unroll
                for (int k = 0; k < MAXC; ++k) {
                    const double* up = nullptr;
                    int uld = 0;
                    if (k < c.nch && !(stage == 2 && !c.cshared[k])) {
                        const int* inv = c.inv + k * ns;
                        const int cc = inv[cs], nbc = c.cnb[k];
                        const int rc = rhs ? 9 * nbc : (inv[rs] >= 0 ? 9 * inv[rs] : -1);
                        if (cc >= 0 && rc >= 0) { uld = 9 * nbc + 1; up = Ubuf + c.cU[k] + rc + (long long)(9 * cc) * uld; }
                    }
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        tv[u][k][j] = 0.0;
                        if (up != nullptr) {
                            if (rhs) { if (j == 0 && lane < 9) tv[u][k][0] = up[(long long)lane * uld]; }
                            else if (lane + 32 * j < 81) tv[u][k][j] = up[ea[j] + eb[j] * uld];
                        }
                    }
                }
            }
            if (blk0 == warp) PHASE(41);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (!valid[u]) continue;
#pragma unroll
                for (int k = 0; k < MAXC; ++k)
#pragma unroll
                    for (int j = 0; j < 3; ++j) v[u][j] += tv[u][k][j];
                if (rhss[u]) { if (lane < 9) P[i0s[u] + (j0s[u] + lane) * ld] = v[u][0]; }
                else {
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        if (lane + 32 * j < 81) P[(i0s[u] + ea[j]) + (j0s[u] + eb[j]) * ld] = v[u][j];
                }
            }
            if (blk0 == warp) PHASE(42);
        }
    }
    __syncthreads();
    PHASE(2);

    // B. right-looking blocked Cholesky of the panel with look-ahead: while warps 1..15 apply block column jb to the
    // trailing columns, warp 0 updates just the next 9x9 diagonal block, factors and inverts it (the serial part), so
    // the single-warp latency hides behind the bulk update.  sLinv is double-buffered.
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
    for (int jb = 0; jb < np; ++jb) {
        const int c0 = 9 * jb;
        const double* sLi = sLinv + 96 * (jb & 1);
        PHASE(10 + 3 * jb);
        if (tid < 81) Linv[81 * (size_t)c.nodes[jb] + tid] = sLi[tid];
        // 2. rows below the diagonal block: x = Lkk^-1 applied from the right
        for (int i = c0 + 9 + tid; i < Rf; i += F3_THREADS) {
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
        // 3. trailing update; warp 0 takes the next diagonal block (update + Cholesky + inverse), warps 1.. the rest
        const int ncb = np - 1 - jb;
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
                for (int cb = jb + 1; cb < np; ++cb) tasks += 3 * ((Rf - 9 * cb + 3) >> 2);
                for (int t = tid - 32; t < tasks; t += F3_THREADS - 32) {
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
    PHASE(3);

    // C. keep the factor for the back-substitution (global panel has leading dimension Rf)
    for (int idx = tid; idx < Rf * Cf; idx += F3_THREADS) {
        int j = idx / Rf, i = idx - j * Rf;
        Lg[idx] = P[i + j * ld];
    }

#ifdef ISLAM_PHASE_CLOCKS
    __syncthreads();
#endif
    PHASE(4);
    // D. update matrix on the boundary (+ rhs row): U = pass-through - L21 L21^T.
    // 4x4 register tiles over the lower triangle; operands are LDS.128 pairs (columns of the panel are 32-byte aligned).
    if (ub > 1) {
        const double* L21 = P + Cf;              // Cf = 9 np; row quads are 32-byte aligned when Cf % 4 == 0
        const int ntr = (ub + 3) >> 2;           // 4-row / 4-column tiles
        // thread tiles flattened over the lower triangle, row tile major: lanes of a warp share the row operand
        const int ntiles = ntr * (ntr + 1) / 2;
        for (int t = tid; t < ntiles; t += F3_THREADS) {
            int tr = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
            while ((tr + 1) * (tr + 2) / 2 <= t) ++tr;
            while (tr * (tr + 1) / 2 > t) --tr;
            const int tc = t - tr * (tr + 1) / 2;
            const bool active = true;
            const int r0 = 4 * tr, s0 = 4 * tc;
            double acc[4][4];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
            const bool full = (r0 + 3 < ub) && (s0 + 3 < ub) && ((Cf & 3) == 0);
            if (full) {
                const double* pa = L21 + r0;
                const double* pb = L21 + s0;
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
            } else {
                int ri[4], si[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { ri[q] = min(r0 + q, ub - 1); si[q] = min(s0 + q, ub - 1); }
                for (int k = 0; k < Cf; ++k) {
                    const double* col = L21 + k * ld;
                    double av[4], bv[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) { av[q] = col[ri[q]]; bv[q] = col[si[q]]; }
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) acc[x][y] += av[x] * bv[y];
                }
            }
            if (!active) continue;
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = -acc[x][y];
            if (stage == 2) {
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        int r = r0 + x, s_ = s0 + y;
                        if (r < ub && s_ < ub && r >= s_) acc[x][y] += base[(Cf + r) + (long long)(Cf + s_) * Rf];
                    }
            }
            for (int k = 0; k < c.nch; ++k) {
                if (stage == 2 && !c.cshared[k]) continue;
                const int* inv = c.inv + k * ns;
                const int nbc = c.cnb[k], ldu = 9 * nbc + 1;
                const double* Uc = Ubuf + c.cU[k];
                int rc[4], cc[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    int r = r0 + q, s_ = s0 + q;
                    if (r >= ub) rc[q] = -1;
                    else if (r == ub - 1) rc[q] = 9 * nbc;
                    else { int t2 = inv[np + r / 9]; rc[q] = t2 >= 0 ? 9 * t2 + r % 9 : -1; }
                    if (s_ >= ub - 1) cc[q] = -1;
                    else { int t2 = inv[np + s_ / 9]; cc[q] = t2 >= 0 ? 9 * t2 + s_ % 9 : -1; }
                }
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y)
                        if (rc[x] >= 0 && cc[y] >= 0 && rc[x] >= cc[y]) acc[x][y] += Uc[rc[x] + (long long)cc[y] * ldu];
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    int r = r0 + x, s_ = s0 + y;
                    if (r < ub && s_ < ub && r >= s_ && !(r == ub - 1 && s_ == ub - 1)) Ug[r + (long long)s_ * ub] = acc[x][y];
                }
        }
    }
#ifdef ISLAM_PHASE_CLOCKS
    __syncthreads();
#endif
    PHASE(5);
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
// x_p = L11^-T (y_p - L21^T x_b), blocked by pose: the inverse 9x9 diagonal blocks were stored by the factor
// kernel, so every block step is a 9x9 mat-vec followed by a 9-deep update of the earlier unknowns.
// The whole panel is pulled into shared memory with one burst of independent loads first: the factor was written a
// few hundred MB of traffic ago, so every access is a DRAM-latency access and must not sit on a dependent chain.
constexpr int BS_THREADS2 = 512;
__global__ void __launch_bounds__(BS_THREADS2)
k_backsolve_level(const LMState* __restrict__ st, const int* __restrict__ fronts, FrontMeta m,
                  const double* __restrict__ Lbuf, const double* __restrict__ Linv, double* __restrict__ D, int force,
                  int smem_doubles) {
    const int f = fronts[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    constexpr int NW = BS_THREADS2 / 32;
    extern __shared__ double smem[];
    const int np = m.np[f], nb = m.nb[f];
    const int Cf = 9 * np, Rb = 9 * nb, Rf = Cf + Rb + 1, ld = Rf;
    const int* nodes = m.nodes + m.nodes_off[f];
    const double* Lg = Lbuf + m.Loff[f];
    double* xb = smem;                 // [Rb]
    double* ts = xb + Rb;              // [Cf]
    double* xs = ts + Cf;              // [16]
    double* sLi = xs + 16;             // [np][81] inverse diagonal blocks
    double* sP = sLi + 81 * np;        // [Rf x Cf] panel copy (if it fits)
    const bool staged = (Rb + Cf + 16 + 81 * np + Rf * Cf <= smem_doubles);
    cudaGridDependencySynchronize();           // PDL: the parents' solution (and, for the root, the factor) is complete
    cudaTriggerProgrammaticLaunchCompletion();
    if (!force && !st->active) return;
    for (int r = tid; r < Rb; r += BS_THREADS2) xb[r] = D[9 * (size_t)nodes[np + r / 9] + (r % 9)];
    for (int i = tid; i < 81 * np; i += BS_THREADS2) sLi[i] = Linv[81 * (size_t)nodes[i / 81] + (i % 81)];
    if (staged)
        for (int i = tid; i < Rf * Cf; i += BS_THREADS2) sP[i] = Lg[i];
    const double* Lp = staged ? sP : Lg;
    __syncthreads();
    // ts[c] = y[c] - sum_r L21[r,c] xb[r]      (one warp per column)
    for (int c = w; c < Cf; c += NW) {
        const double* col = Lp + c * ld + Cf;
        double s = 0.0;
        for (int r = lane; r < Rb; r += 32) s += col[r] * xb[r];
        s = warp_sum(s);
        if (lane == 0) ts[c] = col[Rb] - s;       // rhs row holds y = L11^-1 (b - ...)
    }
    __syncthreads();
    for (int jb = np - 1; jb >= 0; --jb) {
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
        for (int k = tid; k < c0; k += BS_THREADS2) {   // ts[k] -= sum_{c in blk} L[c,k] x[c]   (column k, rows c0..c0+8)
            const double* col = Lp + k * ld + c0;
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < 9; ++q) s += col[q] * xs[q];
            ts[k] -= s;
        }
        __syncthreads();
    }
    for (int c = tid; c < Cf; c += BS_THREADS2) D[9 * (size_t)nodes[c / 9] + (c % 9)] = ts[c];
}

}  // namespace islam
