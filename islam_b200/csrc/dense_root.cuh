// Dense root of the elimination tree: the Schur complement onto the loop-closure endpoints.
//
// Loop-closure edges (|i-j| > band_max) make their endpoints the last poses to be eliminated (symbolic.cpp).  With a
// handful of closures the root is an ordinary front; with thousands (BASELINE config 4: 2 000 closures => ~4 000 poses,
// 24 519 unknowns) it is THE dense contraction of this path (north_star: "the dense Schur-complement GEMM where it
// genuinely is a contraction").  It is stored as one column-major panel — lower triangle plus the rhs row, like every
// other front, leading dimension rounded to even — and factored by a blocked right-looking float64 Cholesky in 128-column
// block steps: two 64-column panels (k_root_potrf: diagonal block AND its inverse; k_root_trsm: the panel below as a
// tensor-core product with that inverse; a narrow update of the second panel in between), then ONE K = 128 update of the
// trailing matrix (k_root_syrk: DMMA m8n8k4, operands staged by TMA bulk copies, two CTAs per SM) — the n^3/3 of the
// work.  The back-substitution walks the blocks in reverse on all SMs (k_root_back).  Tiles sit on the absolute 128-grid of
// the matrix, so a tile column has one owner for the whole factorisation: on several GPUs rank r factors / updates the
// tile columns tc with tc % G == r and the factored blocks are broadcast (islam_b200/dist.py, include/islam_pvgo.h).
// Only the tau / phi variables of a closure pose are promoted (symbolic3.cpp): 6 unknowns per pose, not 9.
// Children that share a root variable are coloured on the host and extend-add colour by colour with plain additions: the
// summation order is fixed, so the dense root is bitwise reproducible like the rest of the library (float64 atomics only
// remain as the fall-back for more than 64 colours, and in k_root_orig, whose entries are written once).
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

// multi-GPU: zero the tile columns of other ranks (this rank never reads its stale copy of them), so that a factored block
// can travel as an all-reduce(SUM) of [owner's columns, zeros elsewhere] — NVSwitch reduces and multicasts in the fabric,
// which measured faster than NCCL's broadcast for these 25 MB messages
__global__ void __launch_bounds__(256)
k_root_zero_foreign(const LMState* __restrict__ st, RootView rv, int G, int rank) {
    if (!st->active) return;
    const int c = blockIdx.x;
    if ((c / 128) % G == rank) return;
    double2* col = reinterpret_cast<double2*>(rv.R + (size_t)c * rv.ld);          // ld is even, R 16-byte aligned
    for (int i = threadIdx.x; i < rv.ld / 2; i += 256) col[i] = make_double2(0.0, 0.0);
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
                int want_part, int first, int atomic) {
    // atomic == 0: the children of this launch are one colour class (pvgo.cu): no two of them touch the same root entry
    if (!force && !st->active) return;
    const int k = rv.children[first + blockIdx.x];
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
            double* dst = &rv.R[pr + (size_t)pc * rv.ld];
            if (atomic) atomicAdd(dst, col[r]); else *dst += col[r];
        }
    }
}

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Cholesky of the nbk x nbk diagonal block at k0 AND its inverse, one CTA.  The block is factored in 8-column steps on a
// 128 x 64 shared-memory matrix [A; I]: every row below a step's 8 x 8 diagonal block is multiplied by that block's inverse
// transpose and the trailing columns are updated — applied to the identity rows this yields E = L^-T for free (row r of E
// only becomes non-zero at step r / 8, so the extra work is about half of the factorisation's).  L goes to the lower triangle,
// the strictly upper part of E (E_rc = (L^-1)_cr, c > r) to the unused UPPER triangle of the same block; diag(E) = 1 / diag(L).
// With the explicit inverse the triangular solve of the panel below becomes a DMMA product (k_root_trsm) and the block step
// of the back-substitution a 64 x 64 mat-vec: round 1's substitution loops took 65 + 56 us per panel, 46 ms of config 4's try
// and its whole critical path on several GPUs.  The 8 x 8 diagonal blocks are factored and inverted in registers, redundantly
// by every thread (no hand-over).
constexpr int PO_LD = DR_NB + 1;
constexpr size_t PO_SMEM = sizeof(double) * 2 * DR_NB * PO_LD;

__global__ void __launch_bounds__(256)
k_root_potrf(const LMState* __restrict__ st, RootView rv, int k0, int nbk, int force, int* chol_fail) {
    if (!force && !st->active) return;
    extern __shared__ __align__(16) double po_sm[];
    double (*M)[PO_LD] = reinterpret_cast<double (*)[PO_LD]>(po_sm);      // rows 0..63: A, rows 64..127: E
    const int tid = threadIdx.x;
    {
        double v[16];                                       // all 16 loads of a thread in flight together
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int idx = tid + 256 * q, j = idx >> 6, i = idx & 63;
            v[q] = (i >= j && i < nbk) ? rv.R[(k0 + i) + (size_t)(k0 + j) * rv.ld] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int idx = tid + 256 * q, j = idx >> 6, i = idx & 63;
            M[i][j] = v[q];
            M[DR_NB + i][j] = i == j ? 1.0 : 0.0;
        }
    }
    __syncthreads();
    bool ok = true;
    const int tx = tid & 63, ty = tid >> 6;
    for (int j0 = 0; j0 < nbk; j0 += 8) {
        // (a) the 8 x 8 diagonal block and its inverse, in registers (columns past nbk: identity)
        double l[8][8], wv[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) l[i][j] = (j0 + i < nbk) ? M[j0 + i][j0 + j] : (i == j ? 1.0 : 0.0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            double d = l[k][k];
            if (!(d > 0.0) || !(d < 1e300)) { ok = false; d = 1.0; }
            const double inv = rsqrt(d);
            l[k][k] = d * inv;
            wv[k][k] = inv;
#pragma unroll
            for (int i = k + 1; i < 8; ++i) l[i][k] *= inv;
#pragma unroll
            for (int j = k + 1; j < 8; ++j)
#pragma unroll
                for (int i = j; i < 8; ++i) l[i][j] -= l[i][k] * l[j][k];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = j + 1; i < 8; ++i) {
                double sacc = 0.0;
#pragma unroll
                for (int k = j; k < i; ++k) sacc += l[i][k] * wv[k][j];
                wv[i][j] = -wv[i][i] * sacc;
            }
        const int m = nbk - j0 - 8 > 0 ? nbk - j0 - 8 : 0;              // A rows below the block
        // (b) rows below (A) and the E rows that are non-zero so far: X <- X W^T.  One thread per row.
        if (ty == 0) {
            const int nrows = m + (j0 + 8 < DR_NB ? j0 + 8 : DR_NB);
            if (tx < nrows) {
                double* row = tx < m ? M[j0 + 8 + tx] : M[DR_NB + (tx - m)];
                double x[8], y[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) x[k] = (j0 + k < nbk) ? row[j0 + k] : 0.0;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    y[c] = 0.0;
#pragma unroll
                    for (int k = 0; k <= c; ++k) y[c] += x[k] * wv[c][k];
                }
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (j0 + c < nbk) row[j0 + c] = y[c];
            }
        }
        __syncthreads();
        // the factored diagonal block itself (nobody reads it any more: every thread holds it in registers)
        if (ty == 1 && tx < 36) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j)
                    if (tx == i * (i + 1) / 2 + j && j0 + i < nbk) M[j0 + i][j0 + j] = l[i][j];
        }
        // (c) trailing columns: thread (tx, ty) owns row tx of the row list above and columns j0 + 8 + ty, + 4, ...
        {
            const int nrows = m + (j0 + 8 < DR_NB ? j0 + 8 : DR_NB);
            if (tx < nrows && m > 0) {
                const bool isA = tx < m;
                const int ri = isA ? j0 + 8 + tx : DR_NB + (tx - m);
                double x[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) x[k] = M[ri][j0 + k];
                for (int j = j0 + 8 + ty; j < nbk; j += 4) {
                    if (isA && ri < j) continue;                      // upper triangle of A
                    double sacc = 0.0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) sacc += x[k] * M[j][j0 + k];
                    M[ri][j] -= sacc;
                }
            }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < DR_NB * DR_NB; idx += 256) {
        const int j = idx >> 6, i = idx & 63;
        if (i >= nbk || j >= nbk) continue;
        rv.R[(k0 + i) + (size_t)(k0 + j) * rv.ld] = (i >= j) ? M[i][j] : M[DR_NB + i][j];
    }
    if (!ok && tid == 0) *chol_fail = 1;
}

// rows below the diagonal block (including the rhs row): X = A L_kk^-T = A E on the tensor cores, 128 rows per CTA.
// E (upper triangular, see k_root_potrf) is the "B" operand [k][col]; the rows are read completely before they are overwritten.
constexpr int TR_LDA = 128 + 4, TR_LDB = DR_NB + 4;
constexpr size_t TR_SMEM = sizeof(double) * DR_NB * (TR_LDA + TR_LDB);
__global__ void __launch_bounds__(128)
k_root_trsm(const LMState* __restrict__ st, RootView rv, int k0, int nbk, int force) {
    if (!force && !st->active) return;
    extern __shared__ __align__(16) double tr_sm[];
    double* As = tr_sm;                          // [k][row]
    double* Bs = tr_sm + DR_NB * TR_LDA;         // [k][col] = E[k][col]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int r0 = k0 + nbk + blockIdx.x * 128;
    const int nbk4 = (nbk + 3) & ~3;
    for (int idx = tid; idx < 128 * nbk4; idx += 128) {
        const int k = idx >> 7, r = idx & 127;
        const bool va = k < nbk && r0 + r <= rv.n;
        const double* ga = rv.R + (va ? (size_t)(r0 + r) + (size_t)(k0 + k) * rv.ld : 0);
        const unsigned sa = (unsigned)__cvta_generic_to_shared(As + k * TR_LDA + r);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(ga), "r"(va ? 8 : 0));
    }
    asm volatile("cp.async.commit_group;");
    {
        double v[32];                                       // k fastest: column c of the block is contiguous
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            const int idx = tid + 128 * q, c = idx >> 6, k = idx & 63;
            v[q] = (c < nbk && k <= c) ? rv.R[(k0 + k) + (size_t)(k0 + c) * rv.ld] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            const int idx = tid + 128 * q, c = idx >> 6, k = idx & 63;
            Bs[k * TR_LDB + c] = (k == c && c < nbk) ? 1.0 / v[q] : v[q];
        }
    }
    asm volatile("cp.async.wait_group 0;");
    __syncthreads();
    const int wr = 32 * w, lk = lane & 3, lm = lane >> 2;
    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int k4 = 0; k4 < nbk4; k4 += 4) {
        const double* ap = As + (k4 + lk) * TR_LDA + wr + lm;
        const double* bp = Bs + (k4 + lk) * TR_LDB + lm;
        double a[4], b[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = ap[8 * i];
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
                const int c = 8 * j + 2 * lk + e;
                if (c < nbk) rv.R[r + (size_t)(k0 + c) * rv.ld] = acc[i][j][e];
            }
    }
}

// Trailing update C -= A_i A_j^T on the FP64 tensor cores: mma.sync m8n8k4 (DMMA).  This is the one genuinely dense
// contraction of the path (north_star: "tensor cores only on the dense Schur-complement GEMM"); on B200 DMMA has the peak of
// the DFMA pipe (tools/pipe_bench: 63 of 64 FMA/clk/SM), and one warp instruction carries 256 multiply-adds, so operand
// traffic from shared memory is a non-issue.  What bounded round 1's kernel (18 TFLOP/s) was arithmetic intensity and
// serialisation: K = 64 per pass re-reads and re-writes the whole 2.4 GB trailing matrix 383 times, and one 256-thread CTA per
// SM (139 KB of operands) first loads, then computes, then stores.  Now:
//   * K = 128 per pass: two 64-column panels are factored (the second after a NARROW update of just its own 64 columns)
//     before the wide update, which halves the trailing-matrix traffic;
//   * 128-thread CTAs own 128 x 64 tiles, K streamed in chunks of 32 through a 2-stage cp.async pipeline (2 x 51 KB), TWO
//     CTAs per SM: while one CTA waits for its C tile or drains its stores, the other one issues DMMA;
//   * tiles sit on the ABSOLUTE 128-grid of the matrix (entries left of / above `base` are masked), so that a tile column
//     has one owner for the whole factorisation: with G ranks, rank r updates the tile columns tc with tc % G == r (1-D
//     block-column-cyclic; G = 1 on a single GPU).  The grid enumerates this rank's tiles only.
constexpr int SY_T = 128;                       // tile-column width = ownership granule = columns factored per block step
constexpr int SY_TN = 64;                       // columns of one CTA's tile
constexpr int SY_KC = 32;                       // k chunk
constexpr int SY_LDA = SY_T + 4, SY_LDB = SY_TN + 4;      // row strides = 4 mod 16 doubles: conflict-free fragment loads
constexpr int SY_STAGE = SY_KC * (SY_LDA + SY_LDB);       // doubles per pipeline stage
constexpr size_t SY_SMEM = sizeof(double) * 2 * SY_STAGE;
constexpr int SY_WN = 32;                       // warp tile width: 32 -> 8 warps per CTA, 16 per SM (64 -> 4 / 8)
constexpr int SY_THREADS = 128 * (SY_TN / SY_WN);

// this rank's 128 x 128 tiles (each one = two CTAs) of the update of columns [base, c_hi)
__host__ __device__ inline long long root_syrk_tiles(int n, int base, int c_hi, int G, int rank, int* T_out, int* j0_out,
                                                     int* Tc_out) {
    const int tc0 = base / SY_T;                                 // first tile column / row with anything to update
    const int T = (n + 1 + SY_T - 1) / SY_T - tc0;               // tile rows (down to the rhs row) below / at tc0
    const int Tc = ((c_hi < n ? c_hi : n) + SY_T - 1) / SY_T - tc0;      // tile columns
    const int j0 = ((rank - tc0) % G + G) % G;                   // first owned column, relative to tc0
    if (T_out) *T_out = T;
    if (j0_out) *j0_out = j0;
    if (Tc_out) *Tc_out = Tc;
    long long total = 0;
    for (int j = j0; j < Tc; j += G) total += T - j;
    return total;
}

// one arrival + the byte count of ALL bulk copies of a stage, then the copies themselves (they only complete_tx)
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_copy_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Operand chunks come in through the TMA engine: row k of a chunk is one contiguous column segment of the factor (128
// doubles for A, 64 for B; ld is even and tiles start on even rows, so every segment is 16-byte aligned), i.e. 64 bulk
// copies per chunk issued by ONE thread and counted on the stage's mbarrier.  (The 8-byte cp.async version spent 87 % of its
// issue slots on copy instructions and their addresses and stalled on the LSU queues: ncu r02_root_syrk_*.)  nk is 64 or
// 128 in every launch (the ragged last block of the matrix has nothing to update), so chunks are always complete.
template <int WN>                                   // warp tile 32 x WN; (SY_TN / WN) * 4 warps per CTA
__global__ void __launch_bounds__(128 * (SY_TN / WN), 2)
k_root_syrk(const LMState* __restrict__ st, RootView rv, int k0, int nk, int base, int c_hi, int force, int G, int rank) {
    if (!force && !st->active) return;
    extern __shared__ __align__(16) double sm_syrk[];
    __shared__ __align__(8) unsigned long long full[2];
    int T, tj, Tc;
    root_syrk_tiles(rv.n, base, c_hi, G, rank, &T, &tj, &Tc);
    int t = blockIdx.x >> 1;
    const int half = blockIdx.x & 1;
    while (t >= T - tj) { t -= T - tj; tj += G; }          // at most (tile columns / G) steps
    const int ti = tj + t;
    const int tile0 = (base / SY_T) * SY_T;
    const int r0 = tile0 + SY_T * ti, c0 = tile0 + SY_T * tj + SY_TN * half;
    const int cmax = c_hi < rv.n ? c_hi : rv.n;             // columns [base, cmax)
    if (c0 + SY_TN <= base || c0 >= cmax || r0 + SY_T <= c0) return;       // nothing of this tile is in the update
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    constexpr int NJ = WN / 8;
    const int wr = 32 * (w & 3), wc = WN * (w >> 2);        // this warp's 32 x WN piece of the tile
    // entirely above the diagonal / above or left of `base`
    const bool idle = r0 + wr + 32 <= c0 + wc || r0 + wr + 32 <= base || c0 + wc + WN <= base;
    const int lk = lane & 3, lm = lane >> 2;
    const int nchunk = nk / SY_KC;
    const int rowsA = rv.ld - r0 < SY_T ? rv.ld - r0 : SY_T, rowsB = rv.ld - c0 < SY_TN ? rv.ld - c0 : SY_TN;      // even
    auto stage_in = [&](int ch) {                           // thread 0
        double* As = sm_syrk + (ch & 1) * SY_STAGE;          // [k][row]
        double* Bs = As + SY_KC * SY_LDA;                    // [k][col]
        unsigned long long* bar = &full[ch & 1];
        mbar_expect_tx(bar, (unsigned)(SY_KC * (rowsA + rowsB) * sizeof(double)));
        const double* src = rv.R + (size_t)(k0 + ch * SY_KC) * rv.ld;
        for (int k = 0; k < SY_KC; ++k, src += rv.ld) {
            tma_copy_1d(As + k * SY_LDA, src + r0, (unsigned)(rowsA * sizeof(double)), bar);
            tma_copy_1d(Bs + k * SY_LDB, src + c0, (unsigned)(rowsB * sizeof(double)), bar);
        }
    };
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        stage_in(0);
        if (nchunk > 1) stage_in(1);
    }
    // the accumulators start as the C tile itself (its loads overlap the panel copies) and the A fragments are negated:
    // D = (-A) B + C, so the result is stored without a dependent read-modify-write at the end
    const bool interior = c0 >= base && c0 + SY_TN <= cmax && r0 >= c0 + SY_TN - 1 && r0 + SY_T <= rv.n + 1;
    double acc[4][NJ][2];
    double* Cw = rv.R + (size_t)(r0 + wr + lm) + (size_t)(c0 + wc + 2 * lk) * rv.ld;       // this lane's first element
    if (interior) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) acc[i][j][e] = Cw[8 * i + (size_t)(8 * j + e) * rv.ld];
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + wr + 8 * i + lm;
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = c0 + wc + 8 * j + 2 * lk + e;
                    acc[i][j][e] = (!idle && r <= rv.n && c < cmax && r >= c && c >= base) ? Cw[8 * i + (size_t)(8 * j + e) * rv.ld] : 0.0;
                }
        }
    }
    __syncthreads();                                        // the barriers are initialised
    for (int ch = 0; ch < nchunk; ++ch) {
        mbar_wait(&full[ch & 1], (ch >> 1) & 1);
        if (!idle) {
            const double* As = sm_syrk + (ch & 1) * SY_STAGE;
            const double* Bs = As + SY_KC * SY_LDA;
#pragma unroll 2
            for (int k4 = 0; k4 < SY_KC; k4 += 4) {
                const double* ap = As + (k4 + lk) * SY_LDA + wr + lm;
                const double* bp = Bs + (k4 + lk) * SY_LDB + wc + lm;
                double a[4], b[NJ];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = -ap[8 * i];
#pragma unroll
                for (int j = 0; j < NJ; ++j) b[j] = bp[8 * j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        }
        __syncthreads();                                    // everybody is done with this stage: refill it
        if (tid == 0 && ch + 2 < nchunk) stage_in(ch + 2);
    }
    if (idle) return;
    if (interior) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) Cw[8 * i + (size_t)(8 * j + e) * rv.ld] = acc[i][j][e];
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + wr + 8 * i + lm;
        if (r > rv.n) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = c0 + wc + 8 * j + 2 * lk + e;
                if (c < cmax && r >= c && c >= base) Cw[8 * i + (size_t)(8 * j + e) * rv.ld] = acc[i][j][e];
            }
    }
}

// Backward substitution L^T x = y, right-looking over the 64-column blocks from the last to the first.  t starts as y (the
// rhs row of the factor).  The launch for block b: every CTA solves the 64 x 64 triangle L_bb^T x_b = t_b for itself (the
// block is 32 KB out of L2; solving it redundantly costs nothing and saves a grid-wide hand-over), CTA 0 publishes x_b,
// and then the CTAs share the columns c < k0 left of the block: t_c -= L[b rows, c] . x_b, one warp per column (the 64
// rows of a column are 512 contiguous bytes: one coalesced 16-byte load per lane).  All 148 SMs stream the factor once;
// the single-CTA version of round 1 took 180 ms of config 4's 490 ms try for these 2.4 GB.
constexpr int RB_THREADS = 256;
__global__ void k_root_back_init(const LMState* __restrict__ st, RootView rv, double* __restrict__ t, int force) {
    if (!force && !st->active) return;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < rv.n) t[c] = rv.R[rv.n + (size_t)c * rv.ld];
}

__global__ void __launch_bounds__(RB_THREADS)
k_root_back(const LMState* __restrict__ st, RootView rv, int k0, int nbk, double* __restrict__ x, double* __restrict__ t,
            int force) {
    if (!force && !st->active) return;
    __shared__ double xs[DR_NB], ts[DR_NB];
    __shared__ double L[DR_NB][DR_NB + 1];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    {
        double v[16];                                       // the whole block: L below, E = L^-T above the diagonal
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int idx = tid + RB_THREADS * q, j = idx >> 6, i = idx & 63;
            v[q] = (i < nbk && j < nbk) ? rv.R[(k0 + i) + (size_t)(k0 + j) * rv.ld] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int idx = tid + RB_THREADS * q, j = idx >> 6, i = idx & 63;
            L[i][j] = v[q];
        }
    }
    if (tid < DR_NB) ts[tid] = tid < nbk ? t[k0 + tid] : 0.0;
    __syncthreads();
    // x_b = L_bb^-T t_b = E t_b: x_c = t_c / L_cc + sum_{k > c} E_ck t_k   (4 threads per row)
    {
        const int c = tid >> 2, q = tid & 3;
        double sacc = 0.0;
        if (c < nbk)
            for (int k = c + 1 + q; k < nbk; k += 4) sacc += L[c][k] * ts[k];
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
        if (q == 0) xs[c] = c < nbk ? ts[c] / L[c][c] + sacc : 0.0;
    }
    __syncthreads();
    if (blockIdx.x == 0)
        for (int c = tid; c < nbk; c += RB_THREADS) x[k0 + c] = xs[c];
    // columns left of the block; rows k0 .. k0 + 63 of column c (rows past nbk only exist for the last, ragged block)
    const double x0 = xs[2 * lane], x1 = xs[2 * lane + 1];            // zero beyond nbk
    const bool two = 2 * lane + 1 < nbk, one = 2 * lane < nbk;
    for (int c = blockIdx.x * (RB_THREADS / 32) + w; c < k0; c += gridDim.x * (RB_THREADS / 32)) {
        const double* col = rv.R + (size_t)c * rv.ld + k0 + 2 * lane;
        double a0 = 0.0, a1 = 0.0;
        if (two && (((size_t)c * rv.ld + k0) & 1) == 0) { const double2 v = *reinterpret_cast<const double2*>(col); a0 = v.x; a1 = v.y; }
        else { if (one) a0 = col[0]; if (two) a1 = col[1]; }
        double s = a0 * x0 + a1 * x1;
        s = warp_sum(s);
        if (lane == 0) t[c] -= s;
    }
}

__global__ void k_root_scatter(const LMState* __restrict__ st, RootView rv, const double* __restrict__ x, double* __restrict__ D,
                               int force) {
    if (!force && !st->active) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rv.n) D[3 * (size_t)rv.vars[i / 3] + (i % 3)] = x[i];
}

}  // namespace islam
