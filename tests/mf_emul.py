"""Test helper: NumPy emulation of the multifrontal numeric phase driven by the C++ symbolic plan
(islam_plan_* of include/islam_pvgo.h; csrc/symbolic3.cpp).  It mirrors k_factor3 / k_backsolve3 (csrc/solver3.cuh) front
by front with dense NumPy blocks, so the ordering over 3-dof variables, the elimination tree, the original-entry lists and
the push maps can be validated on a machine without a GPU."""
import ctypes as C
import numpy as np

from islam_b200 import _lib


def blocks_from_dense(H, plan, N):
    """Hd (N,9,9), Ho (P,9,9) with Ho[p] = H[lo dofs, hi dofs]."""
    Hd = np.stack([H[9 * n:9 * n + 9, 9 * n:9 * n + 9] for n in range(N)])
    lo, hi = plan['pair_lo'], plan['pair_hi']
    Ho = np.stack([H[9 * a:9 * a + 9, 9 * b:9 * b + 9] for a, b in zip(lo, hi)]) if len(lo) else np.zeros((0, 9, 9))
    return Hd, Ho


_V3 = ['np', 'npad', 'nb', 'vars_off', 'vars', 'Loff', 'Uoff', 'Ioff', 'parent', 'level', 'part', 'child_off', 'children',
       'cmap_off', 'cmap', 'orig_off', 'orig_rs', 'orig_cs', 'orig_src', 'level_off', 'level_fronts', 'var_front',
       'var_slot', 'var_pos', 'root_slot', 'scalars']


def get_plan(N, links, **opts):
    L = _lib.lib()
    links = np.ascontiguousarray(links, dtype=np.int64).reshape(-1, 2)
    o = _lib.PvgoOpts()
    for k, v in opts.items():
        setattr(o, k, v)
    h = C.c_void_p()
    _lib.check(L.islam_plan_build(C.byref(h), N, links.shape[0], links.ctypes.data, C.byref(o)), 'islam_plan_build')
    out = {}
    for name in _V3 + ['pair_lo', 'pair_hi']:
        key = name if name.startswith('pair') else 'v3_' + name
        p = C.c_void_p()
        n = L.islam_plan_array(h, key.encode(), C.byref(p))
        assert n >= 0, name
        ct = C.c_int64 if name in ('Loff', 'Uoff', 'Ioff') else C.c_int32
        out[name] = np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(n,)).copy() if n else np.zeros(0, np.int64)
    L.islam_plan_free(h)
    sc = out['scalars']
    out.update(F=int(sc[0]), n_levels=int(sc[1]), dense_root=int(sc[2]), max_rows=int(sc[3]), max_cols=int(sc[4]),
               max_ub=int(sc[5]), root_pivots=int(sc[6]))
    return out


def _front3(plan, f, Hd, Ho, g, scale, lm_min, lm_max, U, with_orig=True, clamp=True, child_filter=lambda c: True):
    """Frontal matrix of front f the way k_factor3 assembles it (push): zero, original 3x3 blocks, children's update
    matrices.  Returns (Fm lower (Rf x Rf), original pivot diagonal)."""
    npad, nb = int(plan['npad'][f]), int(plan['nb'][f])
    np_ = int(plan['np'][f])
    vs = plan['vars'][plan['vars_off'][f]:plan['vars_off'][f + 1]]
    Cf, Rf = 3 * npad, 3 * (npad + nb) + 1
    Fm = np.zeros((Rf, Rf))
    diag = np.zeros(Cf)
    Hdf, Hof = Hd.reshape(-1), Ho.reshape(-1)
    if with_orig:
        for e in range(plan['orig_off'][f], plan['orig_off'][f + 1]):
            rs, cs, src = int(plan['orig_rs'][e]), int(plan['orig_cs'][e]), int(plan['orig_src'][e])
            arr = Hof if (src & 2) else Hdf
            off = src >> 2
            blk = np.array([[arr[off + (9 * c + r if (src & 1) else 9 * r + c)] for c in range(3)] for r in range(3)])
            if rs == cs:
                d = np.diag(blk).copy()
                diag[3 * cs:3 * cs + 3] = d
                blk = np.tril(blk)
                blk[np.arange(3), np.arange(3)] = np.clip(d, lm_min, lm_max) * scale if clamp else 0.0
            Fm[3 * rs:3 * rs + 3, 3 * cs:3 * cs + 3] += blk
        for cs in range(np_):
            Fm[Rf - 1, 3 * cs:3 * cs + 3] = -g.reshape(-1)[3 * vs[cs]:3 * vs[cs] + 3]
        if clamp:
            for cs in range(np_, npad):                       # dummy pivots: identity
                Fm[3 * cs:3 * cs + 3, 3 * cs:3 * cs + 3] = np.eye(3)
    for k in range(plan['child_off'][f], plan['child_off'][f + 1]):
        c = int(plan['children'][k])
        if not child_filter(c):
            continue
        cm = plan['cmap'][plan['cmap_off'][k]:plan['cmap_off'][k + 1]]
        idx = np.concatenate([np.repeat(3 * cm, 3) + np.tile(np.arange(3), len(cm)), [Rf - 1]])
        assert np.all(np.diff(idx) > 0), 'child map must be monotone'
        Fm[np.ix_(idx, idx)] += np.tril(U[c])
    return np.tril(Fm), diag


def _eliminate3(Fm, Cf):
    F11 = Fm[:Cf, :Cf]
    L11 = np.linalg.cholesky(F11 + np.tril(F11, -1).T)
    L21 = np.linalg.solve(L11, Fm[Cf:, :Cf].T).T
    Uf = Fm[Cf:, Cf:] - np.tril(L21 @ L21.T)
    return (L11, L21), Uf


def _backsolve3(plan, fronts, Lp, D):
    Dv = D.reshape(-1, 3)                                    # variable-major view: var u = 3 pose + component
    for f in fronts:
        npad, np_ = int(plan['npad'][f]), int(plan['np'][f])
        vs = plan['vars'][plan['vars_off'][f]:plan['vars_off'][f + 1]]
        L11, L21 = Lp[f]
        xb = Dv[vs[npad:]].reshape(-1)
        t = L21[-1] - L21[:-1].T @ xb
        x = np.linalg.solve(L11.T, t).reshape(npad, 3)
        Dv[vs[:np_]] = x[:np_]
        assert np.abs(x[np_:]).max(initial=0.0) == 0.0      # dummy pivots solve to exact zeros


def solve(plan, Hd, Ho, g, scale, lm_min=1e-4, lm_max=1e32):
    F = plan['F']
    U, Lp = [None] * F, [None] * F
    order = np.argsort(plan['level'], kind='stable')
    for f in order:
        Fm, _ = _front3(plan, f, Hd, Ho, g, scale, lm_min, lm_max, U)
        Lp[f], U[f] = _eliminate3(Fm, 3 * int(plan['npad'][f]))
    D = np.zeros((Hd.shape[0], 9))
    _backsolve3(plan, order[::-1], Lp, D)
    return D


def owners(plan, links):
    """Factor ownership (csrc/pvgo.cu, islam_pvgo_create): the window of a private variable the factor touches (a VO edge
    touches tau/phi of both poses, an IMU pair all six variables of poses i and i+1), else window 0."""
    vp = plan['part'][plan['var_front']].reshape(-1, 3)
    pose = vp[:, :2].max(1)
    edge_own = np.maximum(pose[links[:, 0]], pose[links[:, 1]]) if len(links) else np.zeros(0, int)
    allv = vp.max(1)
    pair_own = np.maximum(allv[:-1], allv[1:])
    return vp, np.where(edge_own >= 0, edge_own, 0), np.where(pair_own >= 0, pair_own, 0)


def solve_sharded(plan, rank, Hd, Ho, g, scale, allreduce, lm_min=1e-4, lm_max=1e32):
    """One rank of the multi-GPU scheme on the variable plan: Hd/Ho/g hold only this rank's factors."""
    F = plan['F']
    part = plan['part']
    U, Lp = [None] * F, [None] * F
    order = np.argsort(plan['level'], kind='stable')
    for f in order:
        if part[f] == rank:
            Fm, _ = _front3(plan, f, Hd, Ho, g, scale, lm_min, lm_max, U)
            Lp[f], U[f] = _eliminate3(Fm, 3 * int(plan['npad'][f]))
    shared = [int(f) for f in order if part[f] < 0]
    bases = [_front3(plan, f, Hd, Ho, g, scale, lm_min, lm_max, U, True, False, lambda c: part[c] == rank) for f in shared]
    buf = np.concatenate([np.concatenate([b[0].ravel(), b[1]]) for b in bases]) if shared else np.zeros(0)
    buf = allreduce(buf)
    off = 0
    for f in shared:
        npad, np_, nb = int(plan['npad'][f]), int(plan['np'][f]), int(plan['nb'][f])
        Cf, Rf = 3 * npad, 3 * (npad + nb) + 1
        Fm = buf[off:off + Rf * Rf].reshape(Rf, Rf).copy()
        diag = buf[off + Rf * Rf:off + Rf * Rf + Cf]
        off += Rf * Rf + Cf
        dd = np.clip(diag, lm_min, lm_max) * scale
        dd[3 * np_:] = 1.0
        Fm[np.arange(Cf), np.arange(Cf)] += dd
        Fc, _ = _front3(plan, f, Hd, Ho, g, scale, lm_min, lm_max, U, False, False, lambda c: part[c] < 0)
        Lp[f], U[f] = _eliminate3(Fm + Fc, Cf)
    D = np.zeros((Hd.shape[0], 9))
    _backsolve3(plan, [f for f in order[::-1] if part[f] < 0 or part[f] == rank], Lp, D)
    return D
