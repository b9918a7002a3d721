// Dense root of the elimination tree: the Schur complement onto the loop-closure endpoints.
//
// Loop-closure edges (|i-j| > band_max) make their endpoints the last poses to be eliminated (symbolic.cpp).  With a
// handful of closures the root is an ordinary front; with thousands (BASELINE config 4: 2 000 closures => ~4 000 poses,
// 36 000 unknowns) it is THE dense contraction of this path (north_star: "the dense Schur-complement GEMM where it
// genuinely is a contraction").  It is stored as one column-major (n+1) x n panel — lower triangle plus the rhs row, like
// every other front — and factored by a tiled right-looking float64 Cholesky: per 64-column block a single-CTA diagonal
// factorisation, a row-parallel triangular solve of the panel below, and a 64x64-tiled SYRK/GEMM update of the trailing
// matrix (the bulk: n^3/3 flops on the fp64 pipes of all 148 SMs).  The back-substitution walks the blocks in reverse.
// Only the tau / phi variables of a closure pose are promoted (symbolic3.cpp): 6 unknowns per pose, not 9.
// Children scatter their update matrices with float64 atomics (the only non-deterministic summation order in the
// library: a closure pose collects contributions from both chain neighbours and every closure it takes part in).
#pragma once
#include "common.cuh"
#include "solver3.cuh"

namespace islam {

constexpr int DR_NB = 64;

struct RootView {
    double* R;            // (n+1) x n column-major, ld = n + 1
    int n, ld, K;         // n = 3 K unknowns of K root variables
    int front;            // front id of the dense root in the plan
    const int* vars;      // [K] root variables in elimination order
    const int* children;  // child-entry indices k (into Front3Meta::children / cmap_off) of the root's children
    int nchildren;
};

// original 3x3 blocks of the root (diagonal blocks clamped + damped, A.4) and the rhs row; one thread per scalar
__global__ void __launch_bounds__(128)
k_root_orig(const LMState* __restrict__ st, RootView rv, Front3Meta m, const double* __restrict__ Hd,
            const double* __restrict__ Ho, const double* __restrict__ g, const islam_lm_params* __restrict__ prm,
            double forced_scale, double lm_min_, double lm_max_, double* __restrict__ diag_out) {
    // diag_out != nullptr (multi-GPU): this rank only holds ITS factors' share of the normal equations, so the diagonal,
    // whose clamp is non-linear, is summed apart (k_root_diag adds it after the all-reduce)
    if (forced_scale == 0.0 && !st->active) return;
    const double scale = forced_scale != 0.0 ? forced_scale : st->diag_scale;
    const double lm_min = forced_scale != 0.0 ? lm_min_ : prm->lm_min, lm_max = forced_scale != 0.0 ? lm_max_ : prm->lm_max;
    const int o0 = m.orig_off[rv.front], no = m.orig_off[rv.front + 1] - o0;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 9 * no) {
        const int e = idx / 9, k = idx - 9 * e, c = k / 3, r = k - 3 * c;
        const int rs = m.orig_rs[o0 + e], cs = m.orig_cs[o0 + e], src = m.orig_src[o0 + e];
        if (rs == cs && r < c) return;
        const double* arr = (src & 2) ? Ho : Hd;
        double v = arr[(size_t)(src >> 2) + ((src & 1) ? 9 * c + r : 9 * r + c)];
        if (rs == cs && r == c) {
            if (diag_out) { atomicAdd(&diag_out[3 * rs + r], v); return; }
            v = fmin(fmax(v, lm_min), lm_max) * scale;
        }
        atomicAdd(&rv.R[(3 * rs + r) + (size_t)(3 * cs + c) * rv.ld], v);
    } else if (idx < 9 * no + rv.n) {
        const int j = idx - 9 * no;
        atomicAdd(&rv.R[rv.n + (size_t)j * rv.ld], -g[3 * (size_t)rv.vars[j / 3] + j % 3]);
    }
}

// multi-GPU: the summed original diagonal, clamped and damped like k_root_orig does on one GPU
__global__ void __launch_bounds__(128)
k_root_diag(const LMState* __restrict__ st, RootView rv, const double* __restrict__ diag, const islam_lm_params* __restrict__ prm) {
    if (!st->active) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < rv.n) rv.R[j + (size_t)j * rv.ld] += fmin(fmax(diag[j], prm->lm_min), prm->lm_max) * st->diag_scale;
}

// extend-add of the children's (packed) update matrices, one CTA per child.  want_part: ROOT_ALL_CHILDREN, or only the
// children of that pose window (multi-GPU: the rank's own before the all-reduce, -1 = the shared separators after it)
constexpr int ROOT_ALL_CHILDREN = -2;
__global__ void __launch_bounds__(256)
k_root_children(const LMState* __restrict__ st, RootView rv, Front3Meta m, const double* __restrict__ Ubuf, int force,
                int want_part) {
    if (!force && !st->active) return;
    const int k = rv.children[blockIdx.x];
    const int c = m.children[k];
    if (want_part != ROOT_ALL_CHILDREN && m.part[c] != want_part) return;
    const int* cm = m.cmap + m.cmap_off[k];
    const int ub = 3 * m.nb[c] + 1;
    const double* U = Ubuf + m.Uoff[c];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int cc = warp; cc < ub - 1; cc += 8) {
        const int pc = 3 * cm[cc / 3] + cc % 3;
        const double* col = U + f3_ucol(cc, ub);
        for (int r = cc + lane; r < ub; r += 32) {
            const int pr = (r == ub - 1) ? rv.n : 3 * cm[r / 3] + r % 3;
            atomicAdd(&rv.R[pr + (size_t)pc * rv.ld], col[r]);
        }
    }
}

// Cholesky of the nbk x nbk diagonal block at k0 (single CTA, shared memory)
__global__ void __launch_bounds__(256)
k_root_potrf(const LMState* __restrict__ st, RootView rv, int k0, int nbk, int force, int* chol_fail) {
    if (!force && !st->active) return;
    __shared__ double A[DR_NB][DR_NB + 1];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < nbk * nbk; idx += 256) {
        int j = idx / nbk, i = idx - j * nbk;
        A[i][j] = (i >= j) ? rv.R[(k0 + i) + (size_t)(k0 + j) * rv.ld] : 0.0;
    }
    __syncthreads();
    bool ok = true;
    for (int k = 0; k < nbk; ++k) {
        double d = A[k][k];
        if (!(d > 0.0) || !(d < 1e300)) { ok = false; d = 1.0; }
        const double inv = 1.0 / sqrt(d);
        __syncthreads();
        for (int i = k + tid; i < nbk; i += 256) A[i][k] = (i == k) ? d * inv : A[i][k] * inv;
        __syncthreads();
        const int m_ = nbk - k - 1;
        for (int idx = tid; idx < m_ * m_; idx += 256) {
            int j = idx / m_, i = idx - j * m_;
            if (i >= j) A[k + 1 + i][k + 1 + j] -= A[k + 1 + i][k] * A[k + 1 + j][k];
        }
        __syncthreads();
    }
    for (int idx = tid; idx < nbk * nbk; idx += 256) {
        int j = idx / nbk, i = idx - j * nbk;
        if (i >= j) rv.R[(k0 + i) + (size_t)(k0 + j) * rv.ld] = A[i][j];
    }
    if (!ok && tid == 0) *chol_fail = 1;
}

// rows below the diagonal block (including the rhs row): X = A L_kk^-T, 128 rows per CTA
__global__ void __launch_bounds__(128)
k_root_trsm(const LMState* __restrict__ st, RootView rv, int k0, int nbk, int force) {
    if (!force && !st->active) return;
    extern __shared__ double sm[];
    double* L = sm;                          // [nbk][nbk+1] row-major lower
    double* X = sm + DR_NB * (DR_NB + 1);    // [nbk][128]   column c of this CTA's rows
    const int tid = threadIdx.x;
    const int row = k0 + nbk + blockIdx.x * 128 + tid;
    for (int idx = tid; idx < nbk * nbk; idx += 128) {
        int j = idx / nbk, i = idx - j * nbk;
        L[i * (DR_NB + 1) + j] = (i >= j) ? rv.R[(k0 + i) + (size_t)(k0 + j) * rv.ld] : 0.0;
    }
    const bool valid = row <= rv.n;
    for (int c = 0; c < nbk; ++c) X[c * 128 + tid] = valid ? rv.R[row + (size_t)(k0 + c) * rv.ld] : 0.0;
    __syncthreads();
    for (int c = 0; c < nbk; ++c) {
        double s = X[c * 128 + tid];
        for (int k = 0; k < c; ++k) s -= X[k * 128 + tid] * L[c * (DR_NB + 1) + k];
        X[c * 128 + tid] = s / L[c * (DR_NB + 1) + c];
    }
    if (valid)
        for (int c = 0; c < nbk; ++c) rv.R[row + (size_t)(k0 + c) * rv.ld] = X[c * 128 + tid];
}

// trailing update C -= A_i A_j^T over 128x128 tiles of the lower triangle (rows up to and including the rhs row) on the
// FP64 tensor cores: mma.sync m8n8k4 (DMMA).  This is the one genuinely dense contraction of the path (north_star: "tensor
// cores only on the dense Schur-complement GEMM"); on B200 the DMMA path has the same peak as the DFMA pipe, but a warp
// instruction carries 256 multiply-adds, so the kernel is no longer bound by shared-memory operand traffic the way the 4x4
// register-tile SIMT version was.  8 warps per CTA, warp tile 32 x 64 = 4 x 8 MMA tiles, both operand panels (k <= 64) staged
// once in shared memory as [k][row] with a row stride of 136 doubles (conflict-free fragment loads).
constexpr int SY_T = 128, SY_LD = SY_T + 8;
constexpr size_t SY_SMEM = sizeof(double) * 2 * DR_NB * SY_LD;

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Tiles sit on the ABSOLUTE 128-grid of the matrix (entries left of / above `base` = k0 + nbk are masked), so that a tile
// column has one owner for the whole factorisation: with G ranks, rank r updates the tile columns tc with tc % G == r
// (1-D block-column-cyclic; G = 1 on a single GPU).  The grid enumerates this rank's tiles only.
__host__ __device__ inline long long root_syrk_tiles(int n, int base, int G, int rank, int* T_out, int* j0_out) {
    const int tc0 = base / SY_T;                                 // first tile column / row with anything to update
    const int T = (n + 1 + SY_T - 1) / SY_T - tc0;               // tile rows (down to the rhs row) below / at tc0
    const int Tc = (n + SY_T - 1) / SY_T - tc0;                  // tile columns
    int j0 = ((rank - tc0) % G + G) % G;                         // first owned column, relative to tc0
    if (T_out) *T_out = T;
    if (j0_out) *j0_out = j0;
    long long total = 0;
    for (int j = j0; j < Tc; j += G) total += T - j;
    return total;
}

__global__ void __launch_bounds__(256)
k_root_syrk(const LMState* __restrict__ st, RootView rv, int k0, int nbk, int force, int G, int rank) {
    if (!force && !st->active) return;
    extern __shared__ __align__(16) double sm_syrk[];
    double* As = sm_syrk;                      // [k][row], row stride SY_LD
    double* Bs = sm_syrk + DR_NB * SY_LD;
    const int base = k0 + nbk;
    int T, tj;
    root_syrk_tiles(rv.n, base, G, rank, &T, &tj);
    int t = blockIdx.x;
    while (t >= T - tj) { t -= T - tj; tj += G; }          // at most (tile columns / G) steps
    const int ti = tj + t;
    const int tile0 = (base / SY_T) * SY_T;
    const int r0 = tile0 + SY_T * ti, c0 = tile0 + SY_T * tj;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nbk4 = (nbk + 3) & ~3;
    // operand panels: asynchronous 8-byte copies global -> shared (no register staging, all of them in flight at once);
    // out-of-range rows / the k padding are zero-filled (src-size 0)
    for (int idx = tid; idx < SY_T * nbk4; idx += 256) {
        const int k = idx >> 7, r = idx & (SY_T - 1);
        const bool va = k < nbk && r0 + r <= rv.n, vb = k < nbk && c0 + r < rv.n;
        const double* ga = rv.R + (va ? (size_t)(r0 + r) + (size_t)(k0 + k) * rv.ld : 0);
        const double* gb = rv.R + (vb ? (size_t)(c0 + r) + (size_t)(k0 + k) * rv.ld : 0);
        const unsigned sa = (unsigned)__cvta_generic_to_shared(As + k * SY_LD + r);
        const unsigned sb = (unsigned)__cvta_generic_to_shared(Bs + k * SY_LD + r);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(ga), "r"(va ? 8 : 0));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sb), "l"(gb), "r"(vb ? 8 : 0));
    }
    asm volatile("cp.async.commit_group;");
    const int wr = 32 * (w & 3), wc = 64 * (w >> 2);       // this warp's 32 x 64 piece of the tile
    // entirely above the diagonal, or entirely left of / above `base` (already final)
    const bool idle = (ti == tj && wr + 31 < wc) || c0 + wc + 64 <= base || r0 + wr + 32 <= base;
    const int lk = lane & 3, lm = lane >> 2;
    // the accumulators start as the C tile itself (its loads overlap the panel copies) and the A fragments are negated:
    // D = (-A) B + C, so the result is stored without a dependent read-modify-write at the end
    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + wr + 8 * i + lm;
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = c0 + wc + 8 * j + 2 * lk + e;
                acc[i][j][e] = (!idle && r <= rv.n && c < rv.n && r >= c && c >= base) ? rv.R[r + (size_t)c * rv.ld] : 0.0;
            }
    }
    asm volatile("cp.async.wait_group 0;");
    __syncthreads();
    if (idle) return;
    for (int k4 = 0; k4 < nbk4; k4 += 4) {
        const double* ap = As + (k4 + lk) * SY_LD + wr + lm;
        const double* bp = Bs + (k4 + lk) * SY_LD + wc + lm;
        double a[4], b[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = -ap[8 * i];
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = bp[8 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + wr + 8 * i + lm;
        if (r > rv.n) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = c0 + wc + 8 * j + 2 * lk + e;
                if (c < rv.n && r >= c && c >= base) rv.R[r + (size_t)c * rv.ld] = acc[i][j][e];
            }
    }
}

// backward substitution of one block: x_blk = L_kk^-T (y_blk - L[below, blk]^T x_below)
__global__ void __launch_bounds__(256)
k_root_back(const LMState* __restrict__ st, RootView rv, int k0, int nbk, double* __restrict__ x, int force) {
    if (!force && !st->active) return;
    __shared__ double t[DR_NB];
    __shared__ double L[DR_NB][DR_NB + 1];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int idx = tid; idx < nbk * nbk; idx += 256) {
        int j = idx / nbk, i = idx - j * nbk;
        L[i][j] = (i >= j) ? rv.R[(k0 + i) + (size_t)(k0 + j) * rv.ld] : 0.0;
    }
    for (int c = w; c < nbk; c += 8) {
        const double* col = rv.R + (size_t)(k0 + c) * rv.ld;
        double s = 0.0;
        for (int r = k0 + nbk + lane; r < rv.n; r += 32) s += col[r] * x[r];
        s = warp_sum(s);
        if (lane == 0) t[c] = col[rv.n] - s;             // rhs row = y
    }
    __syncthreads();
    if (w == 0) {
        for (int c = nbk - 1; c >= 0; --c) {
            double xc = t[c] / L[c][c];
            __syncwarp();
            if (lane == 0) t[c] = xc;
            for (int k = lane; k < c; k += 32) t[k] -= L[c][k] * xc;
            __syncwarp();
        }
    }
    __syncthreads();
    for (int c = tid; c < nbk; c += 256) x[k0 + c] = t[c];
}

__global__ void k_root_scatter(const LMState* __restrict__ st, RootView rv, const double* __restrict__ x, double* __restrict__ D,
                               int force) {
    if (!force && !st->active) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rv.n) D[3 * (size_t)rv.vars[i / 3] + (i % 3)] = x[i];
}

}  // namespace islam
