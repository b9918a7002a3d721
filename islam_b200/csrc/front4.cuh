// Kernel family 2, generation 4 — one multifrontal front per CTA as a producer / consumer pipeline of specialised warps.
//
// Same job and same data layout as k_factor3 (solver3.cuh): the damped Cholesky that replaces `cholesky_ex / cholesky_solve`
// of PyPose's LM.step as configured at /root/reference/pvgo.py:169-171.  What changed is the schedule inside
// the CTA.  k_factor3 alternated between the serial 9x9 pivot blocks (warp 0) and their shadow work with two CTA-wide barriers
// per 9-column step.  Measurements on B200 (tools/lat_bench.cu, tools/chain_bench.cu, ncu) say where such a front loses time:
// a dependent DFMA costs 8.7 clk, but a 64-bit shuffle 30 clk and a shared-memory load 29 clk — and both travel through the
// one shared-memory pipe of the SM, so under the panel / Schur warps' operand traffic (one wavefront per clock at best) every
// shuffle or load on the pivot chain queues behind them.  Hence:
//
//   * 256 threads, ONE CTA per SM: two warps per SM sub-partition, 255 registers per thread.  The working sets live in
//     registers, not in shared memory.
//   * chain warp (warp 0): every lane holds the WHOLE 9x9 diagonal block and factors it redundantly — no shuffle, no load on
//     the serial path (per column: reciprocal seed -> two fma -> fma -> fma); then every lane forward-substitutes its own
//     pivot rows below the block (rows c0+9+lane, c0+41+lane) against the register copy of L_jj, publishes the block column
//     (mbarrier cbar[jb]) and looks ahead: the update of the NEXT block column by this one.  The later block columns of the
//     pivot block are updated one step behind by the trailing warp (pbar[jb]).  The chain warp's only neighbour on its SM
//     sub-partition is the store warp, which issues no fp64 instruction.
//   * panel warps (3): each thread OWNS an 8 x 9 tile of F21 (8 boundary rows of one block column) in registers from the
//     assembly to the moment its block column is final: right-looking updates without any read-modify-write traffic.  When
//     the chain publishes block column jb, the owners of that block column forward-substitute (no inverse needed), write their
//     rows of L21 once (shared memory for the consumers, global memory for the back-substitution), and arrive on xbar[jb].
//   * Schur warps (3): U -= L21 L21^T in 8 x 8 register tiles (one per thread, resident through all steps, 20 shared-memory
//     wavefronts per 2048 fma), touched once at the end: U_global = U_children - acc.
//   * trailing warp: the pivot-block columns the chain's look-ahead leaves out, then the inverse of the diagonal block for the
//     back-substitution kernel (off everybody's critical path).
//
// No CTA-wide barrier after the assembly; the chain never waits for more than the trailing warp's previous step.  Assembly
// (original entries + extend-add of the children's update matrices, fixed order, no atomics) is k_factor3's.
// Shapes: Cf <= 54 pivot columns and <= 104 boundary rows (one 8 x 9 tile per panel thread); other fronts stay on k_factor3.
#pragma once
#include "solver3.cuh"

namespace islam {

constexpr int F4_NT = 256;                          // threads per CTA: 8 warps, two per SM sub-partition, 255 registers each
constexpr int F4_MAX_STEPS = 6;                     // Cf <= 54
constexpr int F4_MAX_NBR = 104;                     // boundary rows (incl. the right-hand side) <= 13 groups of 8
// header doubles in front of the panel (F3_HEAD = 256 of them): [0,8) cbar, [8,16) xbar, [16,24) pbar (mbarriers, one per
// block step), [24,26) tbar (children's update matrices staged by TMA), [32,96) reciprocal diagonals 1/L_cc of every pivot column, [104,136) the F4Ctx, [192,256) staged child maps
constexpr int F4_CBAR = 0, F4_XBAR = 8, F4_PBAR = 16, F4_TBAR = 24, F4_IS = 32, F4_CTX = 104;
constexpr int F4_SE = 24;                            // staged extend-add: elements of each child per thread

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// every calling thread polls (divergent callers)
__device__ __forceinline__ void mbar_wait_thread(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// a converged warp: one lane polls, the others wait at the warp barrier (which also orders the acquire before their reads)
__device__ __forceinline__ void mbar_wait_warp(unsigned long long* bar, unsigned parity) {
    if ((threadIdx.x & 31) == 0) mbar_wait_thread(bar, parity);
    __syncwarp();
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

// Everything a role needs to know about the front; lives in shared memory (the roles are separate non-inlined functions with
// their own register allocation and read the fields on demand)
struct F4Ctx {
    double* Ug;             // update matrix in global memory
    double* Lg;             // factor panel in global memory (leading dimension Rf)
    double* Linv_g;         // inverse diagonal blocks in global memory (81 per block step)
    int* chol_fail;
    int Cf, Rf, ld, ub, nbs, nch, front;
};
// Shared-memory addresses are rebuilt from the dynamic shared-memory symbol inside every role, never carried through the
// context: a pointer loaded from memory is a GENERIC pointer to the compiler, and generic loads of shared memory (LD.E
// instead of LDS) cost several times the latency — measured: the whole front ran 2x slower.
struct F4Smem {
    double* P;              // panel, (Rf x Cf) column-major with leading dimension ld
    double* Uw;             // children's pass-through of the update matrix (valid if nch > 0)
    double* sIS;            // 1 / L_cc of every pivot column
    unsigned long long *cbar, *xbar, *pbar;
};
__device__ __forceinline__ F4Smem f4_smem(const F4Ctx& c) {
    extern __shared__ double smem[];
    F4Smem s;
    s.P = smem + F3_HEAD + (c.Cf & 1);
    s.Uw = smem + F3_HEAD + c.ld * c.Cf + F3_PAD;
    s.sIS = smem + F4_IS;
    s.cbar = reinterpret_cast<unsigned long long*>(smem + F4_CBAR);
    s.xbar = reinterpret_cast<unsigned long long*>(smem + F4_XBAR);
    s.pbar = reinterpret_cast<unsigned long long*>(smem + F4_PBAR);
    return s;
}

// ---- the chain -------------------------------------------------------------------------------------------------------------
// Column C of the Cholesky factorisation of the 9x9 diagonal block, held redundantly by every lane (C a compile-time constant:
// every register array is indexed statically).  Critical path per column: reciprocal seed, e = 1 - d r0, p = e + e^2,
// m = (x r0)(1 + p), fma: no shuffle, no memory.  The reciprocal square root that scales the stored column is off that path.
template <int C> struct F4Diag {
    static __device__ __forceinline__ void run(double (&D)[9][9], double (&isv)[9], bool& ok) {
        const double d = D[C][C];
        if (!(d > 0.0) || !(d < 1e300)) ok = false;        // a failed factor is abandoned by the LM controller (info = 1)
        double r0;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
        const double e1 = fma(-d, r0, 1.0);
        const double p = fma(e1, e1, e1);                  // x / d = (x r0)(1 + e + e^2), |e|^3 < 2^-60
        const double is = f4_rsqrt(d);
#pragma unroll
        for (int i = C + 1; i < 9; ++i) {
            const double pm = D[i][C] * r0;
            const double mm = fma(pm, p, pm);
#pragma unroll
            for (int c2 = C + 1; c2 <= i; ++c2) D[i][c2] = fma(-mm, D[c2][C], D[i][c2]);
        }
#pragma unroll
        for (int i = C; i < 9; ++i) D[i][C] *= is;         // column C of L (i == C: d / sqrt(d))
        isv[C] = is;
        F4Diag<C + 1>::run(D, isv, ok);
    }
};
template <> struct F4Diag<9> {
    static __device__ __forceinline__ void run(double (&)[9][9], double (&)[9], bool&) {}
};

// x <- x L^-T for one row against the register copy of the diagonal block (forward substitution, 1/L_qq = isv[q])
__device__ __forceinline__ void f4_row_solve(double (&x)[9], const double (&D)[9][9], const double (&isv)[9]) {
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        double s = x[q];
#pragma unroll
        for (int k = 0; k < q; ++k) s = fma(-x[k], D[q][k], s);
        x[q] = s * isv[q];
    }
}

__device__ __forceinline__ void f4_chain(const F4Ctx* __restrict__ cp, int lane) {
    const F4Ctx c = *cp;
    const F4Smem sm = f4_smem(c);
    bool ok = true;
    const int ld = c.ld;
    for (int jb = 0; jb < c.nbs; ++jb) {
        const int c0 = 9 * jb, nbelow = c.Cf - c0 - 9;         // pivot rows below this diagonal block
        PHASE(10 + 3 * jb);
        const bool vA = lane < nbelow, vB = lane + 32 < nbelow, hasB = nbelow > 32;
        double* base = sm.P + (size_t)c0 * ld + c0;              // (row c0, column c0)
        double D[9][9], isv[9], a[9], b[9];
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) D[i][j] = base[i + j * ld];        // broadcast loads
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            a[q] = vA ? base[9 + lane + q * ld] : 0.0;
            b[q] = vB ? base[41 + lane + q * ld] : 0.0;
        }
        if (jb == 2) PHASE(8);
        F4Diag<0>::run(D, isv, ok);
        if (jb == 2) PHASE(9);
        f4_row_solve(a, D, isv);
        if (hasB) f4_row_solve(b, D, isv);
        // publish: L_jj (every lane holds the same values), the reciprocal diagonal, this lane's rows of L
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) base[i + j * ld] = D[i][j];
#pragma unroll
        for (int q = 0; q < 9; ++q) sm.sIS[c0 + q] = isv[q];
        if (vA) {
#pragma unroll
            for (int q = 0; q < 9; ++q) base[9 + lane + q * ld] = a[q];
        }
        if (vB) {
#pragma unroll
            for (int q = 0; q < 9; ++q) base[41 + lane + q * ld] = b[q];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.cbar[jb]);
        PHASE(11 + 3 * jb);
        // this lane's rows also go to the global factor (fire and forget); the diagonal block is the trailing warp's to store
        {
            double* gl = c.Lg + c0 + (size_t)c0 * c.Rf;
            if (vA) {
#pragma unroll
                for (int q = 0; q < 9; ++q) gl[9 + lane + (size_t)q * c.Rf] = a[q];
            }
            if (vB) {
#pragma unroll
                for (int q = 0; q < 9; ++q) gl[41 + lane + (size_t)q * c.Rf] = b[q];
            }
        }
        // look-ahead: block column jb + 1 (all rows below) gets this block column's update, C[i][j] -= sum_q L[i][q] L[j][q].
        // Its rows j are the first nine rows of set A: broadcast from shared memory.  The earlier steps' updates of that block
        // column come from the trailing warp, one step behind (pbar); the same wait keeps its read-modify-writes and ours apart.
        if (jb == 2) PHASE(5);
        const int n2 = nbelow < 9 ? nbelow : 9;
        if (n2 > 0) {
#ifndef ISLAM_CHAIN_ONLY
            if (jb >= 1) mbar_wait_warp(&sm.pbar[jb - 1], 0);
#endif
            double sA[9], sB[9];
#pragma unroll
            for (int u = 0; u < 9; ++u) {
                sA[u] = (u < n2 && vA && lane >= u) ? base[9 + lane + (size_t)(9 + u) * ld] : 0.0;
                sB[u] = (u < n2 && vB) ? base[41 + lane + (size_t)(9 + u) * ld] : 0.0;
            }
            if (jb == 2) PHASE(6);
            // rows j of the next block, three pivot columns at a time: 27 independent broadcast loads in flight, then the fma
            // (a load costs 29 clk: left to itself the compiler pairs every load with its fma and pays that latency 81 times)
#pragma unroll
            for (int q0 = 0; q0 < 9; q0 += 3) {
                double l[3][9];
#pragma unroll
                for (int qq = 0; qq < 3; ++qq)
#pragma unroll
                    for (int u = 0; u < 9; ++u) l[qq][u] = base[9 + (u < n2 ? u : 0) + (q0 + qq) * ld];
#pragma unroll
                for (int qq = 0; qq < 3; ++qq)
#pragma unroll
                    for (int u = 0; u < 9; ++u) {
                        sA[u] = fma(-a[q0 + qq], l[qq][u], sA[u]);
                        if (hasB) sB[u] = fma(-b[q0 + qq], l[qq][u], sB[u]);
                    }
            }
            if (jb == 2) PHASE(7);
#pragma unroll
            for (int u = 0; u < 9; ++u) {
                if (u < n2 && vA && lane >= u) base[9 + lane + (size_t)(9 + u) * ld] = sA[u];
                if (u < n2 && vB) base[41 + lane + (size_t)(9 + u) * ld] = sB[u];
            }
        }
        __syncwarp();
        PHASE(12 + 3 * jb);
    }
    PHASE(4);
    if (!ok && lane == 0) *c.chol_fail = 1;
    FRONT_END(c.front);
}

// ---- Schur warps: U -= L21 L21^T in 8 x 8 register tiles, one per thread, resident through all block steps -----------------
__device__ __forceinline__ void f4_schur(const F4Ctx* __restrict__ cp, int st, int nst) {
    const F4Ctx c = *cp;
    const F4Smem sm = f4_smem(c);
    const int ld = c.ld, ub = c.ub, Cf = c.Cf;
    const int nrg = (ub + 7) >> 3, ube = f3_ube(ub);
    // tiles (gr >= tc) enumerated column block by column block: block tc covers columns [8 tc, 8 tc + 8), row groups tc..nrg-1
    auto tiles_before = [&](int tc) { return tc * nrg - (tc * (tc - 1)) / 2; };
    const int ntiles = ub > 1 ? tiles_before(nrg) : 0;
    auto tile_of = [&](int t, int& gr, int& tc) {
        tc = 0;
        while (tc + 1 < nrg && tiles_before(tc + 1) <= t) ++tc;
        gr = tc + (t - tiles_before(tc));
    };
    // rank-(kn) contribution of panel columns [k0, k0 + kn).  Operand loads of the last row group may run past the panel's last
    // row: finite-or-not garbage that only reaches accumulators that are never stored.
    auto schur_tile = [&](int gr, int tc, int k0, int kn, double (&acc)[8][8]) {
        const double* pa = sm.P + Cf + 8 * gr + (size_t)k0 * ld;
        const double* pb = sm.P + Cf + 8 * tc + (size_t)k0 * ld;
#pragma unroll 3
        for (int k = 0; k < kn; ++k) {
            double2 aa[4], bb[4];
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                aa[y] = *reinterpret_cast<const double2*>(pa + k * ld + 2 * y);
                bb[y] = *reinterpret_cast<const double2*>(pb + k * ld + 2 * y);
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    acc[2 * x][2 * y] += aa[x].x * bb[y].x;     acc[2 * x][2 * y + 1] += aa[x].x * bb[y].y;
                    acc[2 * x + 1][2 * y] += aa[x].y * bb[y].x; acc[2 * x + 1][2 * y + 1] += aa[x].y * bb[y].y;
                }
        }
    };
    // U_global[tile] = U[tile] - acc  (U: the children's pass-through in shared memory; nothing for a leaf)
    auto schur_apply = [&](int gr, int tc, const double (&acc)[8][8]) {
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            const int s_ = 8 * tc + y;
            if (s_ >= ub) continue;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const int r0 = 8 * gr + 2 * h;
                if (r0 + 1 < s_ || r0 >= ube) continue;            // the pair (r0, r0+1) lies above column s_ / below the storage
                const int off = f3_ucol(s_, ub) + r0;
                double2 u = c.nch > 0 ? *reinterpret_cast<const double2*>(sm.Uw + off) : make_double2(0.0, 0.0);
                u.x -= acc[2 * h][y]; u.y -= acc[2 * h + 1][y];
                *reinterpret_cast<double2*>(c.Ug + off) = u;
            }
        }
    };
    const bool has_tile = st < ntiles;
    int gr0 = 0, tc0 = 0;
    if (has_tile) tile_of(st, gr0, tc0);
    double sacc[8][8];
#pragma unroll
    for (int x = 0; x < 8; ++x)
#pragma unroll
        for (int y = 0; y < 8; ++y) sacc[x][y] = 0.0;
    for (int jb = 0; jb < c.nbs; ++jb) {
        mbar_wait_warp(&sm.xbar[jb], 0);                // block column jb of L21 is final
        PHASE_BY(48 + 2 * jb, st == 0);
        if (has_tile) schur_tile(gr0, tc0, 9 * jb, 9, sacc);
        PHASE_BY(49 + 2 * jb, st == 0);
    }
    if (has_tile) schur_apply(gr0, tc0, sacc);
    PHASE_BY(60, st == 0);
    for (int t = st + nst; t < ntiles; t += nst) {     // (more tiles than Schur threads: never with <= 104 boundary rows)
        int gr, tc;
        tile_of(t, gr, tc);
        double acc[8][8];
#pragma unroll
        for (int x = 0; x < 8; ++x)
#pragma unroll
            for (int y = 0; y < 8; ++y) acc[x][y] = 0.0;
        schur_tile(gr, tc, 0, Cf, acc);
        schur_apply(gr, tc, acc);
    }
    FRONT_END(c.front);
}

// ---- panel warps: F21 and the right-hand side row, one register-resident 8 x 9 tile per thread ----------------------------
// Warp pw (0..2) holds block columns pw (lanes 0..12) and pw + 3 (lanes 13..25), lane % 13 = group of 8 boundary rows.
__device__ __forceinline__ void f4_panel(const F4Ctx* __restrict__ cp, int pw, int lane) {
    const F4Ctx c = *cp;
    const F4Smem sm = f4_smem(c);
    const int ld = c.ld, Cf = c.Cf, nbs = c.nbs, nbr = c.ub;
    const int slot = lane / 13, rg = lane - 13 * slot, cb = pw + 3 * slot;
    if (slot > 1 || cb >= nbs || 8 * rg >= nbr) return;
    double* P = sm.P;
    double* prow = P + Cf + 8 * rg;                          // this thread's 8 rows (16-byte aligned)
    double T[8][9];
#pragma unroll
    for (int y = 0; y < 9; ++y)
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const double2 v = *reinterpret_cast<const double2*>(prow + (size_t)(9 * cb + y) * ld + 2 * h);
            T[2 * h][y] = v.x; T[2 * h + 1][y] = v.y;
        }
    for (int jb = 0; jb < cb; ++jb) {
        // T -= X[rows, block jb] L11[block cb rows, block jb]^T once block column jb of L21 is final
        const int c0 = 9 * jb;
        mbar_wait_thread(&sm.xbar[jb], 0);
        const double* px = prow + (size_t)c0 * ld;
        const double* pl = P + 9 * cb + (size_t)c0 * ld;
#pragma unroll 3
        for (int q = 0; q < 9; ++q) {
            double xv[8];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const double2 v = *reinterpret_cast<const double2*>(px + q * ld + 2 * h);
                xv[2 * h] = v.x; xv[2 * h + 1] = v.y;
            }
#pragma unroll
            for (int y = 0; y < 9; ++y) {
                const double bv = pl[q * ld + y];
#pragma unroll
                for (int x = 0; x < 8; ++x) T[x][y] = fma(-xv[x], bv, T[x][y]);
            }
        }
    }
    // this thread's block column: wait for the chain, forward-substitute against L_cb,cb (broadcast loads), write L21 once
    {
        const int c0 = 9 * cb;
        mbar_wait_thread(&sm.cbar[cb], 0);
        PHASE_BY(30 + 3 * cb, rg == 0);
        const double* pd = P + c0 + (size_t)c0 * ld;
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const double isq = sm.sIS[c0 + q];
#pragma unroll
            for (int k = 0; k < q; ++k) {
                const double l = pd[q + k * ld];
#pragma unroll
                for (int x = 0; x < 8; ++x) T[x][q] = fma(-T[x][k], l, T[x][q]);
            }
#pragma unroll
            for (int x = 0; x < 8; ++x) T[x][q] *= isq;
        }
        const bool full = 8 * rg + 8 <= nbr;
#pragma unroll
        for (int y = 0; y < 9; ++y) {
            double* dst = prow + (size_t)(c0 + y) * ld;
            if (full) {
#pragma unroll
                for (int h = 0; h < 4; ++h) *reinterpret_cast<double2*>(dst + 2 * h) = make_double2(T[2 * h][y], T[2 * h + 1][y]);
            } else {
#pragma unroll
                for (int x = 0; x < 8; ++x)
                    if (8 * rg + x < nbr) dst[x] = T[x][y];
            }
        }
        mbar_arrive(&sm.xbar[cb]);                     // (count = number of row groups)
        {
            double* gl = c.Lg + Cf + 8 * rg + (size_t)c0 * c.Rf;
#pragma unroll
            for (int y = 0; y < 9; ++y)
#pragma unroll
                for (int x = 0; x < 8; ++x)
                    if (8 * rg + x < nbr) gl[x + (size_t)y * c.Rf] = T[x][y];
        }
        PHASE_BY(31 + 3 * cb, rg == 0);
    }
    FRONT_END(c.front);
}

// ---- trailing warp: the pivot-block columns >= jb + 2 get block column jb's update (block column jb + 1 is the chain's
// look-ahead), one step behind the chain; then the inverse of the diagonal block for the back-substitution kernel ------------
__device__ __forceinline__ void f4_trailing(const F4Ctx* __restrict__ cp, int lane) {
    const F4Ctx c = *cp;
    const F4Smem sm = f4_smem(c);
    const int ld = c.ld, Cf = c.Cf, nbs = c.nbs;
    double* P = sm.P;
    for (int jb = 0; jb < nbs; ++jb) {
        const int c0 = 9 * jb;
        mbar_wait_warp(&sm.cbar[jb], 0);
        // one task per (row i, block column cb <= block of i), 9 entries
        int ntask = 0;
        for (int cb = jb + 2; cb < nbs; ++cb) ntask += Cf - 9 * cb;
        for (int t = lane; t < ntask; t += 32) {
            int cb = jb + 2, rem = t;
            while (rem >= Cf - 9 * cb) { rem -= Cf - 9 * cb; ++cb; }
            const int i = 9 * cb + rem;
            const double* px = P + i + (size_t)c0 * ld;
            const double* pl = P + 9 * cb + (size_t)c0 * ld;
            double* pd = P + i + (size_t)(9 * cb) * ld;
            double acc[9];
#pragma unroll
            for (int y = 0; y < 9; ++y) acc[y] = pd[y * ld];
#pragma unroll 3
            for (int q = 0; q < 9; ++q) {
                const double xv = px[q * ld];
#pragma unroll
                for (int y = 0; y < 9; ++y) acc[y] = fma(-xv, pl[q * ld + y], acc[y]);
            }
#pragma unroll
            for (int y = 0; y < 9; ++y) pd[y * ld] = acc[y];            // (entries above the diagonal of block cb are never read)
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.pbar[jb]);
        // the diagonal block itself -> global factor (lower triangle; nothing above it is ever read)
        for (int idx = lane; idx < 81; idx += 32) {
            const int j = idx / 9, i = idx - 9 * j;
            if (i >= j) c.Lg[c0 + i + (size_t)(c0 + j) * c.Rf] = P[c0 + i + (size_t)(c0 + j) * ld];
        }
        // column `lane` of L_jj^-1 by forward substitution (x = L^-1 e_lane), stored row-major for k_backsolve3
        if (lane < 9) {
            const double* pd = P + c0 + (size_t)c0 * ld;
            double x[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                double s = (q == lane) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < q; ++k) s = fma(-x[k], pd[q + k * ld], s);
                x[q] = s * sm.sIS[c0 + q];
            }
#pragma unroll
            for (int q = 0; q < 9; ++q) c.Linv_g[81 * jb + 9 * q + lane] = x[q];
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
         const islam_lm_params* __restrict__ prm, int smem_doubles) {
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
    const int* vars = m.vars + m.vars_off[f];
    const int k0 = m.child_off[f], nch = m.child_off[f + 1] - k0;
    const int o0 = m.orig_off[f], no = m.orig_off[f + 1] - o0;
    for (int i = tid; i < ld * Cf + F3_PAD; i += NT) smem[F3_HEAD + i] = 0.0;
    if (nch > 0)
        for (int i = tid; i < ulen; i += NT) Uw[i] = 0.0;
    if (tid < F4_MAX_STEPS) { mbar_init(&cbar[tid], 1); mbar_init(&xbar[tid], (Rb + 1 + 7) >> 3); mbar_init(&pbar[tid], 1); }
    unsigned long long* tbar = reinterpret_cast<unsigned long long*>(smem + F4_TBAR);
    if (tid == 0) { mbar_init(&tbar[0], 1); mbar_init(&tbar[1], 1); }
    // Two children whose update matrices fit behind this front in shared memory: they are pulled in by two TMA bulk copies
    // (issued by one thread right after the grid dependency, in flight while the original entries are assembled) and
    // scattered from shared memory; the immutable destination maps are loaded into registers BEFORE the grid dependency.
    bool staged2 = false;
    int nA = 0, nB = 0;
    double *SA = nullptr, *SB = nullptr;
    const double *UgA = nullptr, *UgB = nullptr;
    unsigned short dregA[F4_SE], dregB[F4_SE];
    if (m.dmap != nullptr && nch == 2) {
        const int cA = m.children[k0], cB = m.children[k0 + 1];
        nA = (int)f3_ulen(3 * m.nb[cA] + 1); nB = (int)f3_ulen(3 * m.nb[cB] + 1);
        const int used = F3_HEAD + ld * Cf + F3_PAD + ulen;
        staged2 = nA <= F4_SE * NT && nB <= F4_SE * NT && used + nA + nB <= smem_doubles;
        if (staged2) {
            SA = smem + used; SB = SA + nA;
            UgA = Ubuf + m.Uoff[cA]; UgB = Ubuf + m.Uoff[cB];
            const unsigned short* dA = m.dmap + m.Uoff[cA];
            const unsigned short* dB = m.dmap + m.Uoff[cB];
#pragma unroll
            for (int u = 0; u < F4_SE; ++u) {
                const int e = tid + u * NT;
                dregA[u] = e < nA ? dA[e] : (unsigned short)0xFFFF;
                dregB[u] = e < nB ? dB[e] : (unsigned short)0xFFFF;
            }
        }
    }
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
    if (staged2 && tid == 0) {
        tma_load_1d(SA, UgA, (unsigned)(nA * sizeof(double)), &tbar[0]);
        tma_load_1d(SB, UgB, (unsigned)(nB * sizeof(double)), &tbar[1]);
    }
    __syncthreads();
    PHASE(1);
    if (!pre) assemble_orig();
    __syncthreads();
    PHASE(2);
    // ---- assembly, part 2: extend-add of the children's update matrices (children one after the other: fixed order) ----
    {
        int cm_off = 0, kdone = 0;
        if (staged2) {
            mbar_wait_warp(&tbar[0], 0);
#pragma unroll
            for (int u = 0; u < F4_SE; ++u)
                if (dregA[u] != 0xFFFF) P[dregA[u]] += SA[tid + u * NT];
            __syncthreads();
            mbar_wait_warp(&tbar[1], 0);
#pragma unroll
            for (int u = 0; u < F4_SE; ++u)
                if (dregB[u] != 0xFFFF) P[dregB[u]] += SB[tid + u * NT];
            __syncthreads();
            kdone = 2;
        } else if (m.dmap != nullptr) {
            constexpr int CH = NT >= 512 ? 12 : 18;      // elements of each child per thread, all in flight together
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

    // ---- roles by warp (8 warps, two per SM sub-partition) ---------------------------------------------------------------
    // sub-partition 0: warp 0 chain, warp 4 trailing warp.  Sub-partitions 1..3: a Schur warp (1, 2, 3) and a panel warp (5, 6, 7).
    static_assert(NT == F4_NT, "the role map is written for 8 warps");
    static_assert(sizeof(F4Ctx) <= 8 * (192 - F4_CTX), "F4Ctx must fit the header");
    F4Ctx* cp = reinterpret_cast<F4Ctx*>(smem + F4_CTX);
    if (tid == 0) {
        F4Ctx c;
        c.Ug = Ug; c.Lg = Lg; c.Linv_g = Linv + m.Ioff[f]; c.chol_fail = chol_fail;
        c.Cf = Cf; c.Rf = Rf; c.ld = ld; c.ub = ub; c.nbs = nbs; c.nch = nch; c.front = f;
        *cp = c;
    }
    __syncthreads();
    if (warp == 0) { f4_chain(cp, lane); return; }
#ifdef ISLAM_CHAIN_ONLY
    return;          // developer experiment (tools/phase_clocks.py): time the chain warp with every other warp gone
#endif
    if (warp == 4) f4_trailing(cp, lane);
    else if (warp <= 3) f4_schur(cp, 32 * (warp - 1) + lane, 96);
    else f4_panel(cp, warp - 5, lane);
}

}  // namespace islam
