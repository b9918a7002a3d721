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
//   U : lower triangle of the (3 nb + 1)^2 boundary block (update matrix handed to the parent), packed in column pairs
//       that start at even offsets (f3_ucol below).
// Assembly is "push": the panel is zeroed, the original 3x3 blocks of J^T W J are stored from a per-front list,
// and the children's update matrices are streamed in (coalesced) and added at host-computed destinations, one
// child after the other: a fixed summation order, no atomics, bitwise deterministic.
// Factorisation: 9-column steps; warp 0 runs the serial chain of 9x9 diagonal blocks, warps 1.. do everything else in
// its shadow (row solve, trailing update, store of finished columns, Schur update with register-resident tiles).
#pragma once
#include "common.cuh"

namespace islam {

#ifdef ISLAM_PHASE_CLOCKS
__device__ long long g_phase_clk[64];
__device__ int g_phase_grid = 1;
#define PHASE(n) do { if (blockIdx.x == gridDim.x / 2 && threadIdx.x == 0 && gridDim.x == g_phase_grid) g_phase_clk[n] = clock64(); } while (0)
// same, stamped by an arbitrary thread (front4.cuh: first thread of the panel / Schur warps)
#define PHASE_BY(n, who) do { if (blockIdx.x == gridDim.x / 2 && (who) && gridDim.x == g_phase_grid) g_phase_clk[n] = clock64(); } while (0)
// per front: globaltimer at CTA start, after the grid dependency, at the end; SM id  (tools/level_timeline.py)
__device__ unsigned long long g_front_t[4][8192];
__device__ unsigned long long g_bs_t[4][8192];
#define BS_T(k, f) do { if (threadIdx.x == 0 && (f) < 8192) g_bs_t[k][f] = (k) == 3 ? (unsigned long long)smid() : gtime(); } while (0)
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %smid;" : "=r"(r)); return r; }
#define FRONT_T(k, f) do { if (threadIdx.x == 0 && (f) < 8192) g_front_t[k][f] = (k) == 3 ? (unsigned long long)smid() : gtime(); } while (0)
// end of a front whose warps finish at different times (front4.cuh): the latest stamp wins
#define FRONT_END(f) do { if ((threadIdx.x & 31) == 0 && (f) < 8192) atomicMax(&g_front_t[2][f], gtime()); } while (0)
#else
#define PHASE(n) do { } while (0)
#define PHASE_BY(n, who) do { } while (0)
#define FRONT_T(k, f) do { } while (0)
#define FRONT_END(f) do { } while (0)
#define BS_T(k, f) do { } while (0)
#endif

constexpr int F3_HEAD = 192 + 64;    // doubles in front of the panel: 2 x 96 inverse diagonal blocks (also the stage-1
                                      // diagonal), then 128 ints of staged child maps.  Every byte counts here: a C2
                                      // separator front is 115 328 bytes, and two of them must fit one SM (233 472)
constexpr int F3_CMAP_INTS = 128;
constexpr int F3_PAD = 8;            // zeroed doubles behind the panel: operand loads of edge tiles may run past its last column
constexpr int F3_PRE = 8;            // original entries per thread whose maps are resolved before the grid dependency

__host__ __device__ __forceinline__ int f3_ld(int Rf) { return (Rf + 3) & ~3; }      // 32-byte aligned columns
// Update matrix U ((3nb+1)^2, lower triangle + rhs row): columns are stored in PAIRS (2j, 2j+1), both from row 2j down to
// row ube-1 (ube = ub rounded up to even).  Every column therefore starts at an even offset relative to an even row, so a
// pair of rows (2i, 2i+1) of any column is one aligned 16-byte access: conflict-free vector read-modify-write in the
// Schur update.  ~3 % larger than the tight packing; the extra slots ((2j, 2j+1) and the pad row) are never read.
__host__ __device__ __forceinline__ int f3_ube(int ub) { return (ub + 1) & ~1; }
__host__ __device__ __forceinline__ long long f3_ulen(int ub) { const long long e = f3_ube(ub); return e * (e / 2 + 1); }
// element (r, c), r >= c, sits at f3_ucol(c, ub) + r
__host__ __device__ __forceinline__ int f3_ucol(int c, int ub) {
    const int e = f3_ube(ub), j = c >> 1;
    return 2 * j * (e - j + 1) + (c & 1) * (e - 2 * j) - 2 * j;
}
__host__ __device__ __forceinline__ int f3_uidx(int r, int c, int ub) { return f3_ucol(c, ub) + r; }
__host__ __device__ __forceinline__ long long f3_smem_doubles(int Rf, int Cf, int ub, int mode) {
    return F3_HEAD + (mode >= 1 ? (long long)f3_ld(Rf) * Cf + F3_PAD : 0) + (mode == 2 ? f3_ulen(ub) : 0);
}

// ---- numeric factorisation of one level ------------------------------------------------------------------------------
// stage 0: local front (original entries + all children) -> factor
// stage 1: multi-GPU, shared front BEFORE the all-reduce: partial frontal matrix (this rank's original entries, undamped,
//          with the partial pivot diagonal kept apart + its private children) dumped to `shared`; no factorisation
// stage 2: multi-GPU, shared front AFTER the all-reduce: `shared` + shared children -> factor
// Per shared front the buffer holds [P compact (Rf x Cf)] [U packed] [original pivot diagonal (Cf)]: PyPose's clamp_ acts
// on the fully summed diagonal of J^T W J before damping (A.4), so the diagonal travels separately.
// Linv: per 9-column block step the inverse of its 9x9 diagonal Cholesky block (row-major), for the back-substitution.
// MODE 2: panel and update matrix in shared memory; 1: panel in shared memory, update matrix accumulated in global memory;
// 0: both in global memory (boundary too wide for shared memory: chain separators that see a big loop-closure root).
template <int NT, int MINB, int MODE>
__global__ void __launch_bounds__(NT, MINB)
k_factor3(const LMState* __restrict__ st, const int* __restrict__ fronts, Front3Meta m,
          const double* __restrict__ Hd, const double* __restrict__ Ho, const double* __restrict__ g,
          double* __restrict__ Lbuf, double* __restrict__ Ubuf, double* __restrict__ Linv,
          double* __restrict__ shared, double lm_min_, double lm_max_, double forced_scale, int stage, int pre_ok,
          int trigger_early, int* chol_fail, const islam_lm_params* __restrict__ prm) {
    // Launched with programmatic stream serialisation (PDL): everything up to cudaGridDependencySynchronize() overlaps the
    // tail of the previous level.  That is always the zero fill and the resolution of the immutable maps; with pre_ok
    // (the previous kernel in the stream is another factor level, so the LM state, J^T W J and J^T W r were complete
    // before IT started) also the whole of A1, the original entries.
    constexpr bool P_SMEM = MODE >= 1, U_SMEM = MODE == 2;
    const int f = fronts[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    PHASE(0);
    FRONT_T(0, f); FRONT_T(3, f);
    extern __shared__ double smem[];
    double* sLinv = smem;                      // 2 x 81 (+ pad): double-buffered inverse diagonal blocks
    const int np = m.np[f], npad = m.npad[f], nb = m.nb[f];
    const int Cf = 3 * npad, Rb = 3 * nb, Rf = Cf + Rb + 1, ub = Rb + 1, nbs = npad / 3;
    const int ld = P_SMEM ? f3_ld(Rf) : Rf;
    const int ulen = (int)f3_ulen(ub);
    double* Lg = Lbuf + m.Loff[f];
    double* Ug = Ubuf + m.Uoff[f];
    // the panel starts one double later when Cf is odd, so that the boundary rows (Cf + even) are 16-byte aligned
    double* P = P_SMEM ? smem + F3_HEAD + (Cf & 1) : Lg;
    double* Uw = U_SMEM ? smem + F3_HEAD + ld * Cf + F3_PAD : Ug;
    const int* vars = m.vars + m.vars_off[f];
    const int k0 = m.child_off[f], nch = m.child_off[f + 1] - k0;
    const int o0 = m.orig_off[f], no = m.orig_off[f + 1] - o0;
    if (P_SMEM)
        for (int i = tid; i < ld * Cf + F3_PAD; i += NT) smem[F3_HEAD + i] = 0.0;
    if (U_SMEM)
        for (int i = tid; i < ulen; i += NT) Uw[i] = 0.0;
    // resolve this thread's first F3_PRE original entries (immutable maps -> source offset, destination) and stage the
    // children's boundary -> slot maps, all before the grid dependency: only the value loads remain afterwards
    int osrc[F3_PRE], odst[F3_PRE];
    auto resolve = [&](int idx, int& src_off, int& dst) {
        const int e = idx / 9, k = idx - 9 * e, c = k / 3, r = k - 3 * c;
        const int rs = m.orig_rs[o0 + e], cs = m.orig_cs[o0 + e], src = m.orig_src[o0 + e];
        dst = -1;
        src_off = 0;
        if (rs == cs && r < c) return;                                 // diagonal block: lower triangle only
        src_off = (((src >> 2) + ((src & 1) ? 9 * c + r : 9 * r + c)) << 1) | ((src >> 1) & 1);   // bit 0: Ho
        dst = (3 * rs + r) + (3 * cs + c) * ld;
        if (rs == cs && r == c) dst |= 0x40000000;                     // a pivot diagonal entry
    };
#pragma unroll
    for (int u = 0; u < F3_PRE; ++u) {
        odst[u] = -1; osrc[u] = 0;
        if (stage != 2 && tid + u * NT < 9 * no) resolve(tid + u * NT, osrc[u], odst[u]);
    }
    int* scm = reinterpret_cast<int*>(smem + 192);
    int cm_total = 0;
    for (int k = 0; k < nch; ++k) cm_total += m.nb[m.children[k0 + k]];
    const bool cm_staged = cm_total <= F3_CMAP_INTS;
    if (cm_staged)
        for (int i = tid; i < cm_total; i += NT) scm[i] = m.cmap[m.cmap_off[k0] + i];       // children's maps are contiguous
    double scale = 1.0, lm_min = 0.0, lm_max = 0.0;
    bool active = true;
    auto load_state = [&]() {
        active = !(forced_scale == 0.0 && !st->active);
        scale = forced_scale != 0.0 ? forced_scale : st->diag_scale;
        lm_min = forced_scale != 0.0 ? lm_min_ : prm->lm_min;
        lm_max = forced_scale != 0.0 ? lm_max_ : prm->lm_max;
    };
    // A1. original entries of J^T W J / -J^T W r first touched by this front (disjoint destinations)
    auto put = [&](int d, double val) {
        if (d & 0x40000000) {
            d &= 0x3fffffff;
            if (stage == 1) { smem[d % ld] = val; val = 0.0; }                             // summed over ranks before the clamp
            else val = fmin(fmax(val, lm_min), lm_max) * scale;                            // clamp, then cumulative damping (A.4)
        }
        P[d] = val;
    };
    auto assemble_orig = [&]() {
        {
            double v[F3_PRE];
#pragma unroll
            for (int u = 0; u < F3_PRE; ++u) v[u] = odst[u] >= 0 ? ((osrc[u] & 1) ? Ho : Hd)[osrc[u] >> 1] : 0.0;
#pragma unroll
            for (int u = 0; u < F3_PRE; ++u)
                if (odst[u] >= 0) put(odst[u], v[u]);
        }
        for (int idx = tid + F3_PRE * NT; idx < 9 * no; idx += NT) {   // fronts with more original entries than that
            int so, d;
            resolve(idx, so, d);
            if (d >= 0) put(d, ((so & 1) ? Ho : Hd)[so >> 1]);
        }
        for (int idx = tid; idx < 3 * np; idx += NT) P[(Rf - 1) + idx * ld] = -g[3 * (size_t)vars[idx / 3] + idx % 3];
        for (int idx = 3 * np + tid; idx < Cf; idx += NT) {        // dummy pivots: identity, decoupled
            if (stage == 1) smem[idx] = 0.0;
            else P[idx + idx * ld] = 1.0;
        }
    };
    const bool pre = pre_ok && P_SMEM && stage != 2;
    if (pre) {
        __syncthreads();                       // zero fill complete
        load_state();
        if (active) assemble_orig();
    }
    cudaGridDependencySynchronize();           // previous level (children's U, LM state) complete and visible
    // the next level may start its preamble now, unless its CTAs would have to squeeze in beside this level's (host's
    // call: more fronts in the two levels together than SMs); then it launches when this level has drained
    if (trigger_early) cudaTriggerProgrammaticLaunchCompletion();
    FRONT_T(1, f);
    if (!pre) load_state();
    if (!active) return;
    if (!P_SMEM)
        for (int i = tid; i < Rf * Cf; i += NT) P[i] = 0.0;
    // a leaf whose update matrix stays in global memory writes it once, at the end (U = -L21 L21^T); otherwise it
    // accumulates in place: children's pass-through first, then the Schur contributions
    const bool u_direct = !U_SMEM && nch == 0 && stage == 0;
    if (!U_SMEM && !u_direct)
        for (int i = tid; i < ulen; i += NT) Ug[i] = 0.0;
    __syncthreads();
    PHASE(1);
    if (stage == 2) {
        const double* base = shared + m.shared_off[f];
        for (int idx = tid; idx < Rf * Cf; idx += NT) {
            const int j = idx / Rf, i = idx - j * Rf;
            double v = base[idx];
            if (i == j) v = (j < 3 * np) ? v + fmin(fmax(base[(size_t)Rf * Cf + ulen + j], lm_min), lm_max) * scale : 1.0;
            P[i + j * ld] = v;
        }
        for (int idx = tid; idx < ulen; idx += NT) Uw[idx] = base[(size_t)Rf * Cf + idx];
    } else if (!pre) {
        assemble_orig();
    }
    __syncthreads();
    PHASE(2);

    // A2. extend-add of the children's update matrices: a warp takes four columns of the child's packed lower triangle
    // at a time (up to sixteen independent coalesced loads in flight per lane); children one after the other (fixed order)
    {
        int cm_off = 0, kdone = 0;
        if (MODE == 2 && m.dmap != nullptr && stage == 0) {
            // fast path, children in pairs: the loads of BOTH update matrices (values + host-computed destinations, two
            // coalesced streams each, up to 12 elements per thread) are in flight together, so the pair costs one memory
            // latency; the additions still happen child after child (fixed summation order)
            constexpr int CH = 12;
            for (; kdone + 1 < nch; kdone += 2) {
                const int cA = m.children[k0 + kdone], cB = m.children[k0 + kdone + 1];
                const int nA = (int)f3_ulen(3 * m.nb[cA] + 1), nB = (int)f3_ulen(3 * m.nb[cB] + 1);
                if (nA > CH * NT || nB > CH * NT) break;
                const double* UA = Ubuf + m.Uoff[cA];
                const double* UB = Ubuf + m.Uoff[cB];
                const unsigned short* dA = m.dmap + m.Uoff[cA];
                const unsigned short* dB = m.dmap + m.Uoff[cB];
                double va[CH], vb[CH];
                int da[CH], db[CH];
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    const int e = tid + u * NT;
                    va[u] = e < nA ? UA[e] : 0.0; da[u] = e < nA ? (int)dA[e] : 0xFFFF;
                    vb[u] = e < nB ? UB[e] : 0.0; db[u] = e < nB ? (int)dB[e] : 0xFFFF;
                }
#pragma unroll
                for (int u = 0; u < CH; ++u)
                    if (da[u] != 0xFFFF) P[da[u]] += va[u];
                __syncthreads();
#pragma unroll
                for (int u = 0; u < CH; ++u)
                    if (db[u] != 0xFFFF) P[db[u]] += vb[u];
                __syncthreads();
                cm_off += m.nb[cA] + m.nb[cB];
            }
        }
        for (int k = kdone; k < nch; ++k) {
            const int c = m.children[k0 + k];
            const int nbc = m.nb[c], ubc = 3 * nbc + 1;
            const int* cm = cm_staged ? scm + cm_off : m.cmap + m.cmap_off[k0 + k];
            cm_off += nbc;
            if (stage == 1 && m.part[c] != m.mypart) continue;         // this rank's private children only
            if (stage == 2 && m.part[c] >= 0) continue;                // shared children only
            const double* Uc = Ubuf + m.Uoff[c];
            if (MODE == 2 && m.dmap != nullptr) {
                // destinations precomputed on the host (offsets into [P | U] in shared memory, same packed order as the
                // child's U): two coalesced streams, eight independent element pairs in flight per thread
                const unsigned short* dm = m.dmap + m.Uoff[c];
                const int n = (int)f3_ulen(ubc);                       // unused slots of the layout carry 0xFFFF
                for (int e0 = tid; e0 < n; e0 += 8 * NT) {
                    double v[8];
                    int d[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int e = e0 + u * NT;
                        v[u] = e < n ? Uc[e] : 0.0;
                        d[u] = e < n ? (int)dm[e] : 0xFFFF;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (d[u] != 0xFFFF) P[d[u]] += v[u];
                }
            } else
            for (int cc0 = 4 * warp; cc0 < ubc - 1; cc0 += 4 * NW) {   // the last column is the unused (rhs, rhs) corner
                for (int r0 = 0; cc0 + r0 < ubc; r0 += 128) {
                    double v[4][4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int cc = cc0 + j;
                        const double* col = Uc + f3_ucol(cc < ubc ? cc : 0, ubc);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int r = cc + r0 + lane + 32 * u;
                            v[j][u] = (cc < ubc - 1 && r < ubc) ? col[r] : 0.0;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int cc = cc0 + j;
                        if (cc >= ubc - 1) continue;
                        const int pc = 3 * cm[cc / 3] + cc % 3;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int r = cc + r0 + lane + 32 * u;
                            if (r >= ubc) continue;
                            const int pr = (r == ubc - 1) ? Rf - 1 : 3 * cm[r / 3] + r % 3;
                            if (pc < Cf) P[pr + pc * ld] += v[j][u];
                            else Uw[f3_uidx(pr - Cf, pc - Cf, ub)] += v[j][u];
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
    PHASE(3);

    if (stage == 1) {                          // dump the partial frontal matrix for the all-reduce
        double* base = shared + m.shared_off[f];
        for (int idx = tid; idx < Rf * Cf; idx += NT) { const int j = idx / Rf, i = idx - j * Rf; base[idx] = P[i + j * ld]; }
        for (int idx = tid; idx < ulen; idx += NT) base[(size_t)Rf * Cf + idx] = Uw[idx];
        for (int idx = tid; idx < Cf; idx += NT) base[(size_t)Rf * Cf + ulen + idx] = smem[idx];
        return;
    }

    // B. right-looking blocked Cholesky of the panel, 9 columns per step.  The 9x9 diagonal blocks form the serial chain
    // (update, Cholesky, inverse: warp 0); everything else is hidden in its shadow by warps 1..: the trailing update of
    // the panel, the store of the finished block column to global memory, and the rank-9 contribution of that block
    // column to the update matrix  U -= L21 L21^T  (so no separate Schur-complement phase remains).  sLinv is
    // double-buffered.
    bool ok = true;
    // warp 0 only.  Lane r < 9 owns row r of the 9x9 diagonal block jbn: (optionally) the rank-9 update from block column
    // jbn-1, then a right-looking Cholesky with one shuffle per needed element, the rows of L back to the panel, and
    // column `lane` of L^-1 by forward substitution with L streamed from shared memory.  ~35 registers, no spills on
    // the serial chain.
    auto diag_block = [&](int jbn, bool update, double* Lout) {
        const int d0 = 9 * jbn, row = d0 + (lane < 9 ? lane : 0);
        double a[9], linv[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) a[q] = P[row + (d0 + q) * ld];
        if (update) {
            const int cprev = d0 - 9;
            double w[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) w[k] = P[row + (cprev + k) * ld];
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                double s_ = 0.0;
#pragma unroll
                for (int k = 0; k < 9; ++k) s_ += w[k] * P[(d0 + q) + (cprev + k) * ld];
                a[q] -= s_;
            }
        }
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            double d = __shfl_sync(0xffffffffu, a[c], c);
            if (!(d > 0.0) || !(d < 1e300)) { ok = false; d = 1.0; }
            const double inv = rsqrt(d);                       // <= 1 ulp; every lane redundantly
            linv[c] = inv;
            const double lc = (lane == c) ? d * inv : a[c] * inv;
            a[c] = lc;
#pragma unroll
            for (int c2 = c + 1; c2 < 9; ++c2) a[c2] -= lc * __shfl_sync(0xffffffffu, lc, c2);
        }
        if (lane < 9) {
#pragma unroll
            for (int q = 0; q < 9; ++q)
                if (q <= lane) P[row + (d0 + q) * ld] = a[q];
        }
        __syncwarp();
        if (lane < 9) {                                        // x = L^-1 e_lane, axpy form
            double x[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) x[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                x[k] *= linv[k];
#pragma unroll
                for (int i = k + 1; i < 9; ++i) x[i] -= P[(d0 + i) + (d0 + k) * ld] * x[k];
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) Lout[9 * i + lane] = x[i];
        }
    };
    // finished block column jb -> global factor (rows from its diagonal block down; nothing above is ever read)
    auto store_block = [&](int jb, int t0, int nt) {      // t0 / nt: first thread and number of threads taking part
        if (!P_SMEM) return;
        const int c0 = 9 * jb, w0 = t0 >> 5, nw = nt >> 5;          // whole warps: one column at a time, lanes along the rows
        for (int q = w0; q < 9; q += nw)
            for (int i = c0 + lane; i < Rf; i += 32) Lg[i + (size_t)(c0 + q) * Rf] = P[i + (c0 + q) * ld];
    };
    // U -= L21[:, blocks] L21[:, blocks]^T over the lower triangle (+ rhs row).  Register tiles of 2 rows x 8 columns with
    // the lanes of a warp along the ROWS: the row operand is one 16-byte load per lane, contiguous across the warp, the
    // eight column operands are four broadcast loads, and the read-modify-write of U is one aligned 16-byte access per
    // column (paired-column layout above) — no bank-conflict replays anywhere.  Tiles are enumerated column block by
    // column block: block tc covers columns [8 tc, 8 tc + 8) and the row pairs from 4 tc on.
    const int npair = f3_ube(ub) >> 1, ncblk = (ub + 7) >> 3;
    auto tiles_before = [&](int tc) { return tc * npair - 2 * tc * (tc - 1); };    // sum_{j<tc} (npair - 4 j)
    const int ntiles = ub > 1 ? tiles_before(ncblk) : 0;
    auto tile_of = [&](int t, int& tp, int& tc) {
        tc = 0;
        while (tc + 1 < ncblk && tiles_before(tc + 1) <= t) ++tc;
        tp = 4 * tc + (t - tiles_before(tc));
    };
    int tp_first = 0, tc_first = 0;            // the first tile of this thread in the shadowed steps (threads 32..NT-1)
    if (tid >= 32 && tid - 32 < ntiles) tile_of(tid - 32, tp_first, tc_first);
    // rank-(kn) contribution of panel columns [c0, c0 + kn) to the 2 x 8 tile (tp, tc), accumulated in registers
    auto schur_tile = [&](int tp, int tc, int c0, int kn, double (&acc)[2][8]) {
        const int r0 = 2 * tp, s0 = 8 * tc;
        const double* pa = P + Cf + r0 + c0 * ld;
        const double* pb = P + Cf + s0 + c0 * ld;
        if (P_SMEM) {
#pragma unroll 3
            for (int k = 0; k < kn; ++k) {
                const double2 a = *reinterpret_cast<const double2*>(pa + k * ld);
                double2 b[4];
#pragma unroll
                for (int y = 0; y < 4; ++y) b[y] = *reinterpret_cast<const double2*>(pb + k * ld + 2 * y);
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    acc[0][2 * y] += a.x * b[y].x; acc[0][2 * y + 1] += a.x * b[y].y;
                    acc[1][2 * y] += a.y * b[y].x; acc[1][2 * y + 1] += a.y * b[y].y;
                }
            }
        } else {
            // panel in global memory: scalar operand loads (rows past the panel read as zero)
#pragma unroll 3
            for (int k = 0; k < kn; ++k) {
                const double a0 = pa[k * ld], a1 = (Cf + r0 + 1 < Rf) ? pa[k * ld + 1] : 0.0;
#pragma unroll
                for (int y = 0; y < 8; ++y) {
                    const double b = (Cf + s0 + y < Rf) ? pb[k * ld + y] : 0.0;
                    acc[0][y] += a0 * b; acc[1][y] += a1 * b;
                }
            }
        }
    };
    // U_global[tile] = U[tile] - acc  (U: children's pass-through, in shared or global memory; none for a direct leaf)
    auto schur_apply = [&](int tp, int tc, const double (&acc)[2][8]) {
        const int r0 = 2 * tp, s0 = 8 * tc;
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            const int s_ = s0 + y;
            if (s_ >= ub || r0 + 1 < s_) continue;                     // the pair (r0, r0+1) lies above column s_
            // the finished tile goes straight to the global update matrix (no separate copy of U out of shared memory)
            const int off = f3_ucol(s_, ub) + r0;
            double2 u = u_direct ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2*>(Uw + off);
            u.x -= acc[0][y]; u.y -= acc[1][y];
            *reinterpret_cast<double2*>(Ug + off) = u;
        }
    };
    // Every thread of warps 1.. owns ONE tile whose accumulators stay in registers through all steps (a rank-9
    // contribution per step, in the shadow of the diagonal chain) and touch U once, at the end.  Tiles beyond that (wide
    // boundaries, or the 256- / 128-thread variants) are done in one pass over all pivot columns after the last step.
    const bool has_tile = tid >= 32 && tid - 32 < ntiles;
    double sacc[2][8];
#pragma unroll
    for (int y = 0; y < 8; ++y) { sacc[0][y] = 0.0; sacc[1][y] = 0.0; }
    auto schur_rest = [&]() {
        for (int t = tid - 32 + (NT - 32); t < ntiles; t += NT - 32) {
            int tp, tc;
            tile_of(t, tp, tc);
            double acc[2][8];
#pragma unroll
            for (int y = 0; y < 8; ++y) { acc[0][y] = 0.0; acc[1][y] = 0.0; }
            schur_tile(tp, tc, 0, Cf, acc);
            schur_apply(tp, tc, acc);
        }
    };
    if (warp == 0) diag_block(0, false, sLinv);
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
        const int ncb = nbs - 1 - jb;
        if (ncb == 0) {                        // last block column: nothing left on the serial chain
            store_block(jb, tid, NT);
            if (tid >= 32) {
                if (has_tile) { schur_tile(tp_first, tc_first, c0, 9, sacc); schur_apply(tp_first, tc_first, sacc); }
                schur_rest();
            }
        } else if (warp == 0) {                // the serial chain: next diagonal block (update + Cholesky + inverse)
            diag_block(jb + 1, true, sLinv + 96 * ((jb + 1) & 1));
        } else {
            // trailing update: column block cb only needs rows >= 9 cb (lower trapezoid); the 9 diagonal rows of block
            // jb+1 are warp 0's
            if (P_SMEM) {
                // register tiles of 2 rows x the 9 columns of one block, lanes along the rows: the row operand is one
                // aligned 16-byte load per lane (pairs start at the parity of Cf, see the panel shift above), the nine
                // column operands are broadcast loads, and the read-modify-write is one 16-byte access per column
                const int par = Cf & 1;
                auto first_pair = [&](int cb) { const int rmin = 9 * cb + (cb == jb + 1 ? 9 : 0); return (rmin - par) >> 1; };
                const int last_pair = (Rf - 1 - par) >> 1;
                int tasks = 0;
                for (int cb = jb + 1; cb < nbs; ++cb) tasks += last_pair - first_pair(cb) + 1;
                for (int t = tid - 32; t < tasks; t += NT - 32) {
                    int cb = jb + 1, rem = t;
                    while (rem >= last_pair - first_pair(cb) + 1) { rem -= last_pair - first_pair(cb) + 1; ++cb; }
                    const int i0 = par + 2 * (first_pair(cb) + rem), j0 = 9 * cb;
                    const int rmin = j0 + (cb == jb + 1 ? 9 : 0);
                    const bool v0 = i0 >= rmin, v1 = i0 + 1 < Rf;         // (i0 + 1 >= rmin and i0 < Rf always hold)
                    double acc[2][9];
#pragma unroll
                    for (int y = 0; y < 9; ++y) { acc[0][y] = 0.0; acc[1][y] = 0.0; }
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        const double* col = P + (c0 + q) * ld;
                        const double2 a = *reinterpret_cast<const double2*>(col + i0);
#pragma unroll
                        for (int y = 0; y < 9; ++y) {
                            const double bv = col[j0 + y];
                            acc[0][y] += a.x * bv; acc[1][y] += a.y * bv;
                        }
                    }
                    if (v0 && v1) {
#pragma unroll
                        for (int y = 0; y < 9; ++y) {
                            double2* dst = reinterpret_cast<double2*>(P + i0 + (j0 + y) * ld);
                            double2 u = *dst;
                            u.x -= acc[0][y]; u.y -= acc[1][y];
                            *dst = u;
                        }
                    } else {
#pragma unroll
                        for (int y = 0; y < 9; ++y) {
                            if (v0) P[i0 + (j0 + y) * ld] -= acc[0][y];
                            if (v1) P[i0 + 1 + (j0 + y) * ld] -= acc[1][y];
                        }
                    }
                }
            } else {
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
            store_block(jb, tid - 32, NT - 32);
            if (has_tile) schur_tile(tp_first, tc_first, c0, 9, sacc);
        }
        __syncthreads();
        PHASE(12 + 3 * jb);
    }
    if (!ok && tid == 0) *chol_fail = 1;
    PHASE(4);
#ifdef ISLAM_PHASE_CLOCKS
    __syncthreads();
#endif
    PHASE(6);
    FRONT_T(2, f);
}

// ---- back-substitution of one level (root first) ---------------------------------------------------------------------
// x_p = L11^-T (y_p - L21^T x_b), blocked by 9 columns: the inverse 9x9 diagonal blocks were stored by the factor
// kernel, so every block step is a 9x9 mat-vec followed by a 9-deep update of the earlier unknowns.
// The whole panel is pulled into shared memory with one burst of independent loads first: the factor was written a
// few hundred MB of traffic ago, so every access is a DRAM-latency access and must not sit on a dependent chain.
constexpr int BS3_THREADS = 512;
__host__ __device__ __forceinline__ long long bs3_smem_doubles(int Rf, int Cf, bool staged) {
    const int Rb = Rf - Cf - 1;
    return ((Rb + Cf + 16 + 9LL * Cf + 1) & ~1LL) + (staged ? (long long)Rf * Cf + 2 : 0);      // panel copy 16-byte aligned
}

// ---- TMA (bulk asynchronous copy engine), 1-D form: global -> shared, completion on an mbarrier -------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// true when the phase with the given parity has completed; gives up after ~1 s (the caller then copies by hand)
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, unsigned parity) {
    const long long t0 = clock64();
    unsigned done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && clock64() - t0 > (1LL << 31)) return false;
    }
    return true;
}

// `chained` (top of the tree, all fronts of several levels in ONE launch, parents first in the grid, everything
// co-resident): a front waits for its parent through a per-front completion counter instead of a kernel boundary — the
// front may start when its parent has completed one solve more than itself — so the levels with 1..64 fronts cost one
// flag round trip each instead of a launch + drain each.  n_late: the first n_late CTAs (the top level) stage their factor
// after the grid dependency (the kernel before may still be writing it); everything else prefetches before it.
__global__ void __launch_bounds__(BS3_THREADS)
k_backsolve3(const LMState* __restrict__ st, const int* __restrict__ fronts, Front3Meta m,
             const double* __restrict__ Lbuf, const double* __restrict__ Linv, double* __restrict__ D, int force,
             int smem_doubles, int n_late, int chained, int* __restrict__ count, int dense_root) {
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
    double* sP = smem + ((Rb + Cf + 16 + 81 * nbs + 1) & ~1);      // [Rf x Cf] panel copy (if it fits), 16-byte aligned
    const bool staged = (bs3_smem_doubles(Rf, Cf, true) <= smem_doubles);
    const bool pre_ok = (int)blockIdx.x >= n_late;
    BS_T(0, f); BS_T(3, f);
    // The factor (L, Linv) of every front below the top level was complete before the previous kernel started, so it is
    // pulled into shared memory while the parents are still solving: the panel by ONE bulk asynchronous copy of the TMA
    // engine (global -> shared, completion counted on an mbarrier; the panels start 16-byte aligned in the L arena), the
    // small inverse diagonal blocks by ordinary loads.  The boundary variable ids are immutable.
    __shared__ __align__(8) unsigned long long panel_bar;
    const unsigned panel_bytes = (unsigned)(((size_t)Rf * Cf * sizeof(double)) & ~(size_t)15);
    if (tid == 0 && staged) mbar_init(&panel_bar, 1);
    __syncthreads();
    auto stage_factor = [&]() {
        if (staged && tid == 0) tma_load_1d(sP, Lg, panel_bytes, &panel_bar);
        for (int i = tid; i < 81 * nbs; i += BS3_THREADS) sLi[i] = Linv[m.Ioff[f] + i];
    };
    if (pre_ok) stage_factor();
    int bvar[2] = {0, 0};              // this thread's boundary rows r = tid, tid + 512 -> index into D
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int r = tid + u * BS3_THREADS;
        if (r < Rb) bvar[u] = 3 * vars[npad + r / 3] + (r % 3);
    }
    const int parent = chained ? m.parent[f] : -1;
    cudaGridDependencySynchronize();           // PDL: the kernel before (parents' solution / the factor) is complete
    cudaTriggerProgrammaticLaunchCompletion();
    if (!force && !st->active) {
        if (staged && pre_ok) mbar_wait(&panel_bar, 0);     // never leave with a bulk copy into this CTA's shared memory in flight
        return;
    }
    int mine = 0;
    if (chained) {
        if (tid == 0) {
            mine = *(volatile int*)(count + f);
            if (parent >= 0 && parent != dense_root) {
                const long long t0 = clock64();            // the parent is co-resident and earlier in the grid; the bound
                while (*(volatile int*)(count + parent) != mine + 1)          // only guards against a wedged device:
                    if (clock64() - t0 > (1LL << 31)) {                       // report it (info = 3) and stop the LM loop
                        LMState* sw = const_cast<LMState*>(st);
                        sw->info = 3; sw->continual = 0;
                        break;
                    }
            }
            __threadfence();
        }
        __syncthreads();
    }
    BS_T(1, f);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int r = tid + u * BS3_THREADS;
        if (r < Rb) xb[r] = D[bvar[u]];
    }
    for (int r = tid + 2 * BS3_THREADS; r < Rb; r += BS3_THREADS) xb[r] = D[3 * (size_t)vars[npad + r / 3] + (r % 3)];
    if (!pre_ok) stage_factor();
    const double* Lp = staged ? sP : Lg;
    if (staged) {
        if (!mbar_wait(&panel_bar, 0))             // (never seen: the engine answers within microseconds)
            for (int i = tid; i < Rf * Cf; i += BS3_THREADS) sP[i] = Lg[i];
        if (tid == 0 && (((size_t)Rf * Cf) & 1)) sP[Rf * Cf - 1] = Lg[Rf * Cf - 1];     // an odd last double is not part of the bulk copy
    }
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
    if (chained) {                             // publish: the solution is written before the counter moves
        __threadfence();
        __syncthreads();
        if (tid == 0) *(volatile int*)(count + f) = mine + 1;
    }
    BS_T(2, f);
}

}  // namespace islam
