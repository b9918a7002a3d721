// Kernel family 2, generation 4 — one multifrontal front per CTA as a three-stage producer / consumer pipeline.
//
// Same job, same data layout and same arithmetic order per entry as k_factor3 (solver3.cuh): the damped Cholesky that
// replaces `cholesky_ex / cholesky_solve` of PyPose's LM.step as configured at /root/reference/pvgo.py:169-171.
// What changed is the schedule inside the CTA.  k_factor3 alternated between the serial 9x9 pivot blocks (warp 0) and their
// shadow work with two CTA-wide barriers per 9-column step; every step cost max(chain, shadow) + barrier skew, and the chain
// itself waited for the shadow warps' trailing update of the next pivot block.  Here the three kinds of work are decoupled
// and only meet through mbarriers:
//
//   chain warp (warp 0)     factors the whole pivot block F11 (Cf x Cf, Cf <= 63) on its own: every remaining pivot row rides
//                           along in registers (lane l owns rows c0+l and c0+32+l), identity rows ride along too so the inverse
//                           of the 9x9 diagonal block costs nothing, the per-column critical path is shuffle -> reciprocal ->
//                           multiply -> fma (the reciprocal square root that scales the stored column is off that path).  After
//                           block jb: arrive on cbar[jb], then look ahead: the update of the NEXT block column by this one.  The
//                           later block columns of the pivot block are updated by the panel warps (pbar[jb] tells when).
//   panel warps             wait cbar[jb]; row-solve the boundary rows of block column jb with the inverse diagonal block;
//                           arrive on xbar[jb]; right-looking update of the later block columns of F21 (4 x 9 register tiles);
//                           store the finished block column of L (and the inverse block) to global memory.
//   Schur warps             wait xbar[jb]; rank-9 contribution of block column jb to the update matrix
//                           U -= L21 L21^T in 4 x 8 register tiles that stay in registers through all steps and touch U once.
//
// No CTA-wide barrier after the assembly; the chain never waits for anybody.  Assembly (original entries + extend-add of the
// children's update matrices, fixed order, no atomics) is the same code as k_factor3's shared-memory variant.
#pragma once
#include "solver3.cuh"

namespace islam {

// header doubles in front of the panel (F3_HEAD = 256 of them): [0,8) cbar, [8,16) xbar, [16,24) pbar (mbarriers, one per block
// step), [24,112) inverse of diagonal block 0, [112,136) the F4Ctx of the front (read by the role functions on demand) (the inverses of blocks >= 1 live in the unused upper part of the pivot block: rows 0..8
// of their own 9 columns), [192,256) staged child maps (as k_factor3)
constexpr int F4_CBAR = 0, F4_XBAR = 8, F4_PBAR = 16, F4_LINV0 = 24, F4_CTX = 112;
constexpr int F4_MAX_STEPS = 7;                     // Cf <= 63

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One lane polls (a warp-wide try_wait is 32 shared-memory probes per spin: with ~500 waiting threads the probes alone
// congest the shared-memory pipe that the chain warp lives on), the others wait at the warp barrier; __syncwarp orders the
// polling lane's acquire before the other lanes' reads.
__device__ __forceinline__ void mbar_wait_forever(unsigned long long* bar, unsigned parity) {
    if ((threadIdx.x & 31) == 0) {
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
    __syncwarp();
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// 1/d to full double precision from the 20-bit hardware seed and two Newton steps: four dependent FMAs, no branches
__device__ __forceinline__ double f4_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
}

// 1/sqrt(d) from the hardware seed and two Newton steps (no special-case branch: d is a checked positive pivot)
__device__ __forceinline__ double f4_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double h = 0.5 * d;
    double t = fma(-h * y, y, 0.5);
    y = fma(y, t, y);
    t = fma(-h * y, y, 0.5);
    return fma(y, t, y);
}

// One column step of the chain warp's 9-column block, C a compile-time constant so that every register array is indexed
// statically.  a: row c0 + lane, b: row c0 + 32 + lane, e: identity row `lane` (lanes < 9) of the block column.
template <int C> struct F4Col {
    static __device__ __forceinline__ void run(double (&a)[9], double (&b)[9], double (&e)[9], bool hasB, int lane, bool& ok) {
        const double d = __shfl_sync(0xffffffffu, a[C], C);
        if (!(d > 0.0) || !(d < 1e300)) ok = false;            // off the critical path; a failed factor is abandoned (info = 1)
        // multipliers m = x / d on the critical path: 20-bit seed r0, e = 1 - d r0, x / d = (x r0) (1 + e + e^2)  (|e|^3 < 2^-60)
        double r0;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
        const double e1 = fma(-d, r0, 1.0);
        const double pa = a[C] * r0, pb = b[C] * r0, pe = e[C] * r0;
        const double p = fma(e1, e1, e1);
        const double ma = fma(pa, p, pa), mb = fma(pb, p, pb), me = fma(pe, p, pe);
        const double is = f4_rsqrt(d);             // off the path: only scales the stored column
#pragma unroll
        for (int c2 = C + 1; c2 < 9; ++c2) {
            const double t = __shfl_sync(0xffffffffu, a[C], c2);       // unscaled entry (row c0 + c2, column c0 + C)
            a[c2] = fma(-ma, t, a[c2]);
            if (hasB) b[c2] = fma(-mb, t, b[c2]);
            e[c2] = fma(-me, t, e[c2]);
        }
        a[C] *= is;                                // (lane C: d / sqrt(d))
        b[C] *= is;
        e[C] *= is;
        F4Col<C + 1>::run(a, b, e, hasB, lane, ok);
    }
};
template <> struct F4Col<9> {
    static __device__ __forceinline__ void run(double (&)[9], double (&)[9], double (&)[9], bool, int, bool&) {}
};

// Everything a role needs to know about the front (the roles are separate non-inlined functions so that each gets its own
// register allocation: the 4 x 8 Schur tiles and the chain's row sets do not fit beside the assembly's live values)
struct F4Ctx {
    double* P;              // panel in shared memory, (Rf x Cf) column-major with leading dimension ld
    double* Uw;             // children's pass-through of the update matrix in shared memory (valid if nch > 0)
    double* Ug;             // update matrix in global memory
    double* Lg;             // factor panel in global memory (leading dimension Rf)
    double* Linv_g;         // inverse diagonal blocks in global memory (81 per block step)
    double* sL0;            // inverse of diagonal block 0 in shared memory
    unsigned long long *cbar, *xbar, *pbar;
    int* chol_fail;
    int Cf, Rf, ld, ub, nbs, nch, front;
    int n_schur, n_panel;   // threads per role (whole warps)
};

// inverse diagonal block jb, element (i, j): block 0 in the header, the others in rows 0..8 above their own 9 columns
__device__ __forceinline__ double f4_linv(const F4Ctx& c, int jb, int i, int j) {   // c: in shared memory
    return jb == 0 ? c.sL0[9 * i + j] : c.P[i + (size_t)(9 * jb + j) * c.ld];
}

// Look-ahead of the chain warp: the NEXT block column (n2 <= 9 columns, all remaining rows) gets the update of the block
// column just factored, C[i][j] -= sum_q L[i][q] L[j][q].  The rows stay in registers (a: row c0 + lane, b: row c0 + 32 +
// lane), the nine rows j are broadcast from shared memory, all 9 (x 2) accumulators independent.  colp = &P(c0, c0).
template <bool HASB>
__device__ __forceinline__ void f4_lookahead(double* colp, int ld, int n2, int lane, bool vA, bool vB,
                                             const double (&a)[9], const double (&b)[9]) {
    double sA[9], sB[9];
#pragma unroll
    for (int u = 0; u < 9; ++u) {
        const int j = 9 + u;
        sA[u] = (u < n2 && vA && lane >= j) ? colp[lane + (size_t)j * ld] : 0.0;
        sB[u] = (HASB && u < n2 && vB) ? colp[32 + lane + (size_t)j * ld] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 9; ++q) {
#pragma unroll
        for (int u = 0; u < 9; ++u) {
            const double l = colp[9 + (u < n2 ? u : 0) + q * ld];
            sA[u] = fma(-a[q], l, sA[u]);
            if (HASB) sB[u] = fma(-b[q], l, sB[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < 9; ++u) {
        const int j = 9 + u;
        if (u < n2 && vA && lane >= j) colp[lane + (size_t)j * ld] = sA[u];
        if (HASB && u < n2 && vB) colp[32 + lane + (size_t)j * ld] = sB[u];
    }
}

// ---- the chain: right-looking Cholesky of the pivot block inside one warp ------------------------------------------------
__device__ __noinline__ void f4_chain(const F4Ctx* __restrict__ cp, int lane) {
    const F4Ctx& c = *cp;
    bool ok = true;
    const int ld = c.ld;
    for (int jb = 0; jb < c.nbs; ++jb) {
        const int c0 = 9 * jb, nr = c.Cf - c0;
        PHASE(10 + 3 * jb);
        const bool vA = lane < nr, vB = lane + 32 < nr, hasB = nr > 32;
        double* colp = c.P + (size_t)c0 * ld + c0;       // (row c0, column c0)
        double a[9], b[9], e[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            a[q] = vA ? colp[lane + q * ld] : 0.0;
            b[q] = vB ? colp[32 + lane + q * ld] : 0.0;
            e[q] = (q == lane) ? 1.0 : 0.0;
        }
        F4Col<0>::run(a, b, e, hasB, lane, ok);
        // rows of L back to the panel; lane l < 9 holds e[q] = (L^-1)[q][l]
        if (vA) {
#pragma unroll
            for (int q = 0; q < 9; ++q)
                if (lane >= 9 || q <= lane) colp[lane + q * ld] = a[q];
        }
        if (vB) {
#pragma unroll
            for (int q = 0; q < 9; ++q) colp[32 + lane + q * ld] = b[q];
        }
        if (lane < 9) {
            if (jb == 0) {
#pragma unroll
                for (int q = 0; q < 9; ++q) c.sL0[9 * q + lane] = e[q];
            } else {
                double* up = c.P + (size_t)(c0 + lane) * ld;      // rows 0..8 of column c0 + lane: above the pivot block's lower part
#pragma unroll
                for (int q = 0; q < 9; ++q) up[q] = e[q];
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&c.cbar[jb]);          // block column jb of L11 and its inverse diagonal block are published
        PHASE(11 + 3 * jb);
        const int n2 = nr - 9 < 9 ? nr - 9 : 9;
        if (n2 > 0) {
            // block column jb + 1 already has the updates of the steps before this one only when the panel warps are through
            // with step jb - 1 (they update the block columns >= jb + 1 of the pivot block, one step behind the chain); the same
            // wait keeps their read-modify-writes and the look-ahead's apart
#ifndef ISLAM_CHAIN_ONLY
            if (jb >= 1) mbar_wait_forever(&c.pbar[jb - 1], 0);
#endif
            if (hasB) f4_lookahead<true>(colp, ld, n2, lane, vA, vB, a, b);
            else f4_lookahead<false>(colp, ld, n2, lane, vA, vB, a, b);
        }
        __syncwarp();
        PHASE(12 + 3 * jb);
    }
    PHASE(4);
    if (!ok && lane == 0) *c.chol_fail = 1;
    FRONT_END(c.front);
}

// ---- Schur warps: U -= L21 L21^T in 4 x 8 register tiles, one per thread, resident through all block steps -----------------
__device__ __noinline__ void f4_schur(const F4Ctx* __restrict__ cp, int st) {
    const F4Ctx& c = *cp;
    const int ld = c.ld, ub = c.ub, Cf = c.Cf;
    const int ngrp = (ub + 3) >> 2, ncblk = (ub + 7) >> 3, ube = f3_ube(ub);
    // column block tc covers columns [8 tc, 8 tc + 8) of U and the row groups (4 rows) from 2 tc on
    auto tiles_before = [&](int tc) { return tc * ngrp - tc * (tc - 1); };      // sum_{j<tc} (ngrp - 2 j)
    const int ntiles = ub > 1 ? tiles_before(ncblk) : 0;
    auto tile_of = [&](int t, int& gr, int& tc) {
        tc = 0;
        while (tc + 1 < ncblk && tiles_before(tc + 1) <= t) ++tc;
        gr = 2 * tc + (t - tiles_before(tc));
    };
    // rank-(kn) contribution of panel columns [k0, k0 + kn) to the tile (gr, tc).  Operand loads of the last row group /
    // column block may run past the panel's last row: finite-or-not garbage that only reaches accumulators never stored.
    auto schur_tile = [&](int gr, int tc, int k0, int kn, double (&acc)[4][8]) {
        const double* pa = c.P + Cf + 4 * gr + (size_t)k0 * ld;
        const double* pb = c.P + Cf + 8 * tc + (size_t)k0 * ld;
#pragma unroll 3
        for (int k = 0; k < kn; ++k) {
            const double2 a01 = *reinterpret_cast<const double2*>(pa + k * ld);
            const double2 a23 = *reinterpret_cast<const double2*>(pa + k * ld + 2);
            double2 bb[4];
#pragma unroll
            for (int y = 0; y < 4; ++y) bb[y] = *reinterpret_cast<const double2*>(pb + k * ld + 2 * y);
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                acc[0][2 * y] += a01.x * bb[y].x; acc[0][2 * y + 1] += a01.x * bb[y].y;
                acc[1][2 * y] += a01.y * bb[y].x; acc[1][2 * y + 1] += a01.y * bb[y].y;
                acc[2][2 * y] += a23.x * bb[y].x; acc[2][2 * y + 1] += a23.x * bb[y].y;
                acc[3][2 * y] += a23.y * bb[y].x; acc[3][2 * y + 1] += a23.y * bb[y].y;
            }
        }
    };
    // U_global[tile] = U[tile] - acc  (U: the children's pass-through in shared memory; nothing for a leaf)
    auto schur_apply = [&](int gr, int tc, const double (&acc)[4][8]) {
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            const int s_ = 8 * tc + y;
            if (s_ >= ub) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r0 = 4 * gr + 2 * h;
                if (r0 + 1 < s_ || r0 >= ube) continue;            // the pair (r0, r0+1) lies above column s_ / below the storage
                const int off = f3_ucol(s_, ub) + r0;
                double2 u = c.nch > 0 ? *reinterpret_cast<const double2*>(c.Uw + off) : make_double2(0.0, 0.0);
                u.x -= acc[2 * h][y]; u.y -= acc[2 * h + 1][y];
                *reinterpret_cast<double2*>(c.Ug + off) = u;
            }
        }
    };
    const bool has_tile = st < ntiles;
    int gr0 = 0, tc0 = 0;
    if (has_tile) tile_of(st, gr0, tc0);
    double sacc[4][8];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 8; ++y) sacc[x][y] = 0.0;
    for (int jb = 0; jb < c.nbs; ++jb) {
        mbar_wait_forever(&c.xbar[jb], 0);             // block column jb of L21 is final
        PHASE_BY(48 + 2 * jb, st == 0);
        if (has_tile) schur_tile(gr0, tc0, 9 * jb, 9, sacc);
        PHASE_BY(49 + 2 * jb, st == 0);
    }
    if (has_tile) schur_apply(gr0, tc0, sacc);
    PHASE_BY(60, st == 0);
    for (int t = st + c.n_schur; t < ntiles; t += c.n_schur) {   // tiles beyond one per thread (boundaries wider than the CTA)
        int gr, tc;
        tile_of(t, gr, tc);
        double acc[4][8];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 8; ++y) acc[x][y] = 0.0;
        schur_tile(gr, tc, 0, Cf, acc);
        schur_apply(gr, tc, acc);
    }
    FRONT_END(c.front);
}

// ---- panel warps: the boundary rows of L (F21 and the right-hand side) and the trailing part of the pivot block ------------
__device__ __noinline__ void f4_panel(const F4Ctx* __restrict__ cp, int pt) {
    const F4Ctx& c = *cp;
    const int ld = c.ld, Cf = c.Cf, nbs = c.nbs, n_panel = c.n_panel;
    const int nbr = c.ub;                               // boundary rows incl. the right-hand-side row
    const int ngrp = (nbr + 3) >> 2;                    // groups of 4 boundary rows; the last one may be partial
    double* P = c.P;
    for (int jb = 0; jb < nbs; ++jb) {
        const int c0 = 9 * jb;
        mbar_wait_forever(&c.cbar[jb], 0);
        PHASE_BY(30 + 3 * jb, pt == 0);
        // rows below the pivot block: x = a Lkk^-T, one thread per row
        for (int i = pt; i < nbr; i += n_panel) {
            double* row = P + Cf + i + (size_t)c0 * ld;
            double av[9], x[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) av[q] = row[q * ld];
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                double s_ = 0.0;
#pragma unroll
                for (int k = 0; k <= q; ++k) s_ += av[k] * f4_linv(c, jb, q, k);
                x[q] = s_;
            }
#pragma unroll
            for (int q = 0; q < 9; ++q) row[q * ld] = x[q];
        }
        named_bar_sync(2, n_panel);
        if (pt == 0) mbar_arrive(&c.xbar[jb]);
        PHASE_BY(31 + 3 * jb, pt == 0);
        // pivot rows of block columns >= jb + 2 (block column jb + 1 is the chain's look-ahead): one task per
        // (row i, block column cb <= block of i), 9 entries.  First, because the chain waits for it (pbar).
        {
            int ntask = 0;
            for (int cb = jb + 2; cb < nbs; ++cb) ntask += Cf - 9 * cb;
            for (int t = pt; t < ntask; t += n_panel) {
                int cb = jb + 2, rem = t;
                while (rem >= Cf - 9 * cb) { rem -= Cf - 9 * cb; ++cb; }
                const int i = 9 * cb + rem;
                const double* px = P + i + (size_t)c0 * ld;
                const double* pl = P + 9 * cb + (size_t)c0 * ld;
                double acc[9];
#pragma unroll
                for (int y = 0; y < 9; ++y) acc[y] = 0.0;
#pragma unroll 3
                for (int q = 0; q < 9; ++q) {
                    const double xv = px[q * ld];
#pragma unroll
                    for (int y = 0; y < 9; ++y) acc[y] += xv * pl[q * ld + y];
                }
                double* pd = P + i + (size_t)(9 * cb) * ld;
#pragma unroll
                for (int y = 0; y < 9; ++y) pd[y * ld] -= acc[y];       // (entries above the diagonal of block cb are never read)
            }
            named_bar_sync(2, n_panel);
            if (pt == 0) mbar_arrive(&c.pbar[jb]);
        }
        // right-looking update of the later block columns of F21: 4 rows x 9 columns per task
        const int ncb = nbs - 1 - jb;
        for (int t = n_panel - 1 - pt; t < ncb * ngrp; t += n_panel) {
            const int cb = jb + 1 + t / ngrp, rg = t % ngrp;
            const double* px = P + Cf + 4 * rg + (size_t)c0 * ld;         // rows of block column jb
            const double* pl = P + 9 * cb + (size_t)c0 * ld;              // L11[9 cb .. 9 cb + 8][c0 ..]
            double acc[4][9];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 9; ++y) acc[x][y] = 0.0;
#pragma unroll 3
            for (int q = 0; q < 9; ++q) {
                const double2 x01 = *reinterpret_cast<const double2*>(px + q * ld);
                const double2 x23 = *reinterpret_cast<const double2*>(px + q * ld + 2);
#pragma unroll
                for (int y = 0; y < 9; ++y) {
                    const double bv = pl[q * ld + y];
                    acc[0][y] += x01.x * bv; acc[1][y] += x01.y * bv;
                    acc[2][y] += x23.x * bv; acc[3][y] += x23.y * bv;
                }
            }
            double* pd = P + Cf + 4 * rg + (size_t)(9 * cb) * ld;
            if (4 * rg + 4 <= nbr) {
#pragma unroll
                for (int y = 0; y < 9; ++y) {
                    double2* d01 = reinterpret_cast<double2*>(pd + y * ld);
                    double2* d23 = reinterpret_cast<double2*>(pd + y * ld + 2);
                    double2 u = *d01, v = *d23;
                    u.x -= acc[0][y]; u.y -= acc[1][y]; v.x -= acc[2][y]; v.y -= acc[3][y];
                    *d01 = u; *d23 = v;
                }
            } else {                                    // last, partial row group: rows past the panel are nobody's
#pragma unroll
                for (int y = 0; y < 9; ++y)
#pragma unroll
                    for (int x = 0; x < 4; ++x)
                        if (4 * rg + x < nbr) pd[y * ld + x] -= acc[x][y];
            }
        }
        named_bar_sync(2, n_panel);
        PHASE_BY(32 + 3 * jb, pt == 0);
    }
    PHASE_BY(61, pt == 0);
    FRONT_END(c.front);
}

// ---- store warps (the chain warp's neighbours on its SM sub-partition: no fp64 work there): finished block columns of L and the
// inverse diagonal blocks -> global memory ------------------------------------------------------------------------------------
__device__ __noinline__ void f4_store(const F4Ctx* __restrict__ cp, int wt, int nthreads) {
    const F4Ctx& c = *cp;
    const int ld = c.ld, Cf = c.Cf, Rf = c.Rf;
    for (int jb = 0; jb < c.nbs; ++jb) {
        const int c0 = 9 * jb;
        mbar_wait_forever(&c.cbar[jb], 0);             // pivot rows of block column jb (rows from its diagonal block down)
        for (int i = wt; i < 81; i += nthreads) c.Linv_g[81 * jb + i] = f4_linv(c, jb, i / 9, i % 9);
        const int np_ = Cf - c0;
        for (int idx = wt; idx < 9 * np_; idx += nthreads) {
            const int q = idx / np_, i = c0 + idx - q * np_;
            c.Lg[i + (size_t)(c0 + q) * Rf] = c.P[i + (size_t)(c0 + q) * ld];
        }
        mbar_wait_forever(&c.xbar[jb], 0);             // boundary rows
        const int nb_ = Rf - Cf;
        for (int idx = wt; idx < 9 * nb_; idx += nthreads) {
            const int q = idx / nb_, i = Cf + idx - q * nb_;
            c.Lg[i + (size_t)(c0 + q) * Rf] = c.P[i + (size_t)(c0 + q) * ld];
        }
    }
    FRONT_END(c.front);
}

template <int NT>
__global__ void __launch_bounds__(NT, 1)
k_front4(const LMState* __restrict__ st, const int* __restrict__ fronts, Front3Meta m,
         const double* __restrict__ Hd, const double* __restrict__ Ho, const double* __restrict__ g,
         double* __restrict__ Lbuf, double* __restrict__ Ubuf, double* __restrict__ Linv,
         double lm_min_, double lm_max_, double forced_scale, int pre_ok, int* chol_fail,
         const islam_lm_params* __restrict__ prm) {
    const int f = fronts[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    FRONT_T(0, f); FRONT_T(3, f);
    PHASE(0);
    extern __shared__ double smem[];
    const int np = m.np[f], npad = m.npad[f], nb = m.nb[f];
    const int Cf = 3 * npad, Rb = 3 * nb, Rf = Cf + Rb + 1, ub = Rb + 1, nbs = npad / 3;
    const int ld = f3_ld(Rf);
    const int ulen = (int)f3_ulen(ub);
    double* Lg = Lbuf + m.Loff[f];
    double* Ug = Ubuf + m.Uoff[f];
    double* P = smem + F3_HEAD + (Cf & 1);               // boundary rows (Cf + even) 16-byte aligned
    double* Uw = smem + F3_HEAD + ld * Cf + F3_PAD;
    unsigned long long* cbar = reinterpret_cast<unsigned long long*>(smem + F4_CBAR);
    unsigned long long* xbar = reinterpret_cast<unsigned long long*>(smem + F4_XBAR);
    unsigned long long* pbar = reinterpret_cast<unsigned long long*>(smem + F4_PBAR);
    double* sL0 = smem + F4_LINV0;
    const int* vars = m.vars + m.vars_off[f];
    const int k0 = m.child_off[f], nch = m.child_off[f + 1] - k0;
    const int o0 = m.orig_off[f], no = m.orig_off[f + 1] - o0;
    for (int i = tid; i < ld * Cf + F3_PAD; i += NT) smem[F3_HEAD + i] = 0.0;
    if (nch > 0)
        for (int i = tid; i < ulen; i += NT) Uw[i] = 0.0;
    if (tid < F4_MAX_STEPS) { mbar_init(&cbar[tid], 1); mbar_init(&xbar[tid], 1); mbar_init(&pbar[tid], 1); }
    // ---- assembly, part 1 (before the grid dependency): immutable maps -> registers / shared memory ------------------
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
        if (tid + u * NT < 9 * no) resolve(tid + u * NT, osrc[u], odst[u]);
    }
    int* scm = reinterpret_cast<int*>(smem + 192);
    int cm_total = 0;
    for (int k = 0; k < nch; ++k) cm_total += m.nb[m.children[k0 + k]];
    const bool cm_staged = cm_total <= F3_CMAP_INTS;
    if (cm_staged)
        for (int i = tid; i < cm_total; i += NT) scm[i] = m.cmap[m.cmap_off[k0] + i];
    double scale = 1.0, lm_min = 0.0, lm_max = 0.0;
    bool active = true;
    auto load_state = [&]() {
        active = !(forced_scale == 0.0 && !st->active);
        scale = forced_scale != 0.0 ? forced_scale : st->diag_scale;
        lm_min = forced_scale != 0.0 ? lm_min_ : prm->lm_min;
        lm_max = forced_scale != 0.0 ? lm_max_ : prm->lm_max;
    };
    auto put = [&](int d, double val) {
        if (d & 0x40000000) {
            d &= 0x3fffffff;
            val = fmin(fmax(val, lm_min), lm_max) * scale;             // clamp, then cumulative damping (A.4)
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
        for (int idx = tid + F3_PRE * NT; idx < 9 * no; idx += NT) {
            int so, d;
            resolve(idx, so, d);
            if (d >= 0) put(d, ((so & 1) ? Ho : Hd)[so >> 1]);
        }
        for (int idx = tid; idx < 3 * np; idx += NT) P[(Rf - 1) + idx * ld] = -g[3 * (size_t)vars[idx / 3] + idx % 3];
        for (int idx = 3 * np + tid; idx < Cf; idx += NT) P[idx + idx * ld] = 1.0;      // dummy pivots: identity, decoupled
    };
    const bool pre = pre_ok != 0;
    if (pre) {
        __syncthreads();                       // zero fill complete
        load_state();
        if (active) assemble_orig();
    }
    cudaGridDependencySynchronize();           // previous level (children's U, LM state) complete and visible
    cudaTriggerProgrammaticLaunchCompletion();
    FRONT_T(1, f);
    if (!pre) load_state();
    if (!active) return;
    __syncthreads();
    PHASE(1);
    if (!pre) assemble_orig();
    __syncthreads();
    PHASE(2);
    // ---- assembly, part 2: extend-add of the children's update matrices (children one after the other: fixed order) ----
    {
        int cm_off = 0, kdone = 0;
        if (m.dmap != nullptr) {
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
            const double* Uc = Ubuf + m.Uoff[c];
            if (m.dmap != nullptr) {
                const unsigned short* dm = m.dmap + m.Uoff[c];
                const int n = (int)f3_ulen(ubc);
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
            } else {
                for (int cc0 = 4 * warp; cc0 < ubc - 1; cc0 += 4 * NW) {
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
            }
            __syncthreads();
        }
    }
    PHASE(3);

    // ---- the three-stage pipeline: roles by warp -------------------------------------------------------------------------
    // warp 0: chain.  The other warps of its SM sub-partition (4, 8, ..): store warps, which issue no fp64 instruction, so the
    // chain's dependent stream never queues behind somebody else's DFMAs.  The rest: Schur warps (one 4 x 8 tile per thread),
    // then panel warps.
    static_assert(sizeof(F4Ctx) <= 8 * (192 - F4_CTX), "F4Ctx must fit the header");
    F4Ctx* cp = reinterpret_cast<F4Ctx*>(smem + F4_CTX);
    constexpr int N_STORE_W = NW / 4 - 1, N_COMP_W = NW - NW / 4;
    const int ngrp = (ub + 3) >> 2, ncblk = (ub + 7) >> 3;
    const int ntiles = ub > 1 ? ncblk * ngrp - ncblk * (ncblk - 1) : 0;
    int n_schur_w = (ntiles + 31) >> 5;
    if (n_schur_w > (2 * N_COMP_W) / 3) n_schur_w = (2 * N_COMP_W) / 3;
    if (tid == 0) {
        F4Ctx c;
        c.P = P; c.Uw = Uw; c.Ug = Ug; c.Lg = Lg; c.Linv_g = Linv + m.Ioff[f]; c.sL0 = sL0;
        c.cbar = cbar; c.xbar = xbar; c.pbar = pbar; c.chol_fail = chol_fail;
        c.Cf = Cf; c.Rf = Rf; c.ld = ld; c.ub = ub; c.nbs = nbs; c.nch = nch; c.front = f;
        c.n_schur = 32 * n_schur_w;
        c.n_panel = 32 * (N_COMP_W - n_schur_w);
        *cp = c;
    }
    __syncthreads();
    if (warp == 0) { f4_chain(cp, lane); return; }
#ifdef ISLAM_CHAIN_ONLY
    return;          // developer experiment (tools/phase_clocks.py): time the chain warp with every other warp gone
#endif
    if ((warp & 3) == 0) { f4_store(cp, 32 * ((warp >> 2) - 1) + lane, 32 * N_STORE_W); return; }
    const int k = (warp >> 2) * 3 + (warp & 3) - 1;          // index among the compute warps
    if (k < n_schur_w) f4_schur(cp, 32 * k + lane);
    else f4_panel(cp, 32 * (k - n_schur_w) + lane);
}

}  // namespace islam
