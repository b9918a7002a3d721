"""Test helper: NumPy emulation of the multifrontal numeric phase driven by the C++ symbolic plan
(islam_plan_* of include/islam_pvgo.h).  It mirrors k_factor_level / k_backsolve_level front by front with dense
NumPy blocks, so the ordering, elimination tree and gather maps can be validated on a machine without a GPU."""
import ctypes as C
import numpy as np

from islam_b200 import _lib

_I64 = {'f_Loff', 'f_Uoff'}


def get_plan(N, links, **opts):
    L = _lib.lib()
    links = np.ascontiguousarray(links, dtype=np.int64).reshape(-1, 2)
    o = _lib.PvgoOpts()
    for k, v in opts.items():
        setattr(o, k, v)
    h = C.c_void_p()
    _lib.check(L.islam_plan_build(C.byref(h), N, links.shape[0], links.ctypes.data, C.byref(o)), 'islam_plan_build')
    out = {}
    for name in ['pair_lo', 'pair_hi', 'pair_adj', 'pair_eoff', 'pair_edges', 'node_eoff', 'node_edges', 'edge_pair',
                 'f_np', 'f_nb', 'f_nodes_off', 'f_nodes', 'f_Loff', 'f_Uoff', 'f_parent', 'f_level', 'f_part',
                 'f_child_off', 'f_children', 'c_inv_off', 'c_inv', 'f_hmap_off', 'hmap', 'level_off', 'level_fronts',
                 'node_front', 'node_slot', 'node_pos']:
        p = C.c_void_p()
        n = L.islam_plan_array(h, name.encode(), C.byref(p))
        assert n >= 0, name
        ct = C.c_int64 if name in _I64 else C.c_int32
        out[name] = np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(n,)).copy() if n else np.zeros(0, np.int64)
    L.islam_plan_free(h)
    return out


def blocks_from_dense(H, plan, N):
    """Hd (N,9,9), Ho (P,9,9) with Ho[p] = H[lo dofs, hi dofs]."""
    Hd = np.stack([H[9 * n:9 * n + 9, 9 * n:9 * n + 9] for n in range(N)])
    lo, hi = plan['pair_lo'], plan['pair_hi']
    Ho = np.stack([H[9 * a:9 * a + 9, 9 * b:9 * b + 9] for a, b in zip(lo, hi)]) if len(lo) else np.zeros((0, 9, 9))
    return Hd, Ho


def solve(plan, Hd, Ho, g, scale, lm_min=1e-4, lm_max=1e32):
    """Returns D (N,9) solving (H with clamped, damped diagonal) D = -g, the way the CUDA kernels do."""
    F = len(plan['f_np'])
    N = Hd.shape[0]
    nodes_off, nodes = plan['f_nodes_off'], plan['f_nodes']
    U = [None] * F
    Lp = [None] * F
    order = np.argsort(plan['f_level'], kind='stable')
    for f in order:
        np_, nb = int(plan['f_np'][f]), int(plan['f_nb'][f])
        ns = np_ + nb
        nd = nodes[nodes_off[f]:nodes_off[f] + ns]
        Cf, Rf = 9 * np_, 9 * ns + 1
        Fm = np.zeros((Rf, Rf))
        hm = plan['hmap'][plan['f_hmap_off'][f]:plan['f_hmap_off'][f + 1]].reshape(ns, np_)
        for cs in range(np_):
            nc = nd[cs]
            blk = Hd[nc].copy()
            d = np.clip(np.diag(blk), lm_min, lm_max) * scale
            blk[np.arange(9), np.arange(9)] = d
            Fm[9 * cs:9 * cs + 9, 9 * cs:9 * cs + 9] = blk
            Fm[Rf - 1, 9 * cs:9 * cs + 9] = -g[nc]
            for rs in range(cs + 1, ns):
                h = hm[rs, cs]
                if h >= 0:
                    b = Ho[h >> 1]
                    Fm[9 * rs:9 * rs + 9, 9 * cs:9 * cs + 9] = b.T if (h & 1) else b
        for k in range(plan['f_child_off'][f], plan['f_child_off'][f + 1]):
            c = int(plan['f_children'][k])
            inv = plan['c_inv'][plan['c_inv_off'][k]:plan['c_inv_off'][k + 1]]
            nbc = int(plan['f_nb'][c])
            idx = np.full(9 * nbc + 1, -1)
            for s in range(ns):
                if inv[s] >= 0:
                    idx[9 * inv[s]:9 * inv[s] + 9] = np.arange(9 * s, 9 * s + 9)
            idx[9 * nbc] = Rf - 1
            assert (idx >= 0).all(), 'child boundary not contained in parent front'
            assert np.all(np.diff(idx) > 0), 'child map must be monotone'
            Fm[np.ix_(idx, idx)] += np.tril(U[c])
        Fm = np.tril(Fm)
        F11 = Fm[:Cf, :Cf]
        L11 = np.linalg.cholesky(F11 + np.tril(F11, -1).T)
        L21 = np.linalg.solve(L11, Fm[Cf:, :Cf].T).T
        Lp[f] = (L11, L21)
        Uf = Fm[Cf:, Cf:] - np.tril(L21 @ L21.T)
        U[f] = Uf
    D = np.zeros((N, 9))
    for f in order[::-1]:
        np_, nb = int(plan['f_np'][f]), int(plan['f_nb'][f])
        nd = nodes[nodes_off[f]:nodes_off[f] + np_ + nb]
        L11, L21 = Lp[f]
        xb = D[nd[np_:]].reshape(-1)
        y = L21[-1]
        t = y - L21[:-1].T @ xb
        x = np.linalg.solve(L11.T, t)
        D[nd[:np_]] = x.reshape(np_, 9)
    return D


def owners(plan, links):
    """Factor ownership rule of csrc/pvgo.cu (islam_pvgo_create): the window of a private endpoint, else window 0."""
    node_part = plan['f_part'][plan['node_front']]
    a, b = node_part[links[:, 0]], node_part[links[:, 1]]
    edge_owner = np.where(a >= 0, a, np.where(b >= 0, b, 0))
    pa, pb = node_part[:-1], node_part[1:]
    pair_owner = np.where(pa >= 0, pa, np.where(pb >= 0, pb, 0))
    return node_part, edge_owner, pair_owner


def solve_sharded(plan, rank, Hd, Ho, g, scale, allreduce, lm_min=1e-4, lm_max=1e32):
    """One rank of the multi-GPU scheme in NumPy: Hd/Ho/g hold only the contributions of the factors this
    rank owns.  Private fronts are eliminated locally; shared fronts get base = partial originals + private children,
    `allreduce(buffer)` sums it over ranks, then every rank factors the shared fronts and back-substitutes."""
    F = len(plan['f_np'])
    N = Hd.shape[0]
    part = plan['f_part']
    nodes_off, nodes = plan['f_nodes_off'], plan['f_nodes']
    U, Lp = [None] * F, [None] * F
    order = np.argsort(plan['f_level'], kind='stable')

    def front_matrix(f, with_orig, clamp, child_filter):
        np_, nb = int(plan['f_np'][f]), int(plan['f_nb'][f])
        ns = np_ + nb
        nd = nodes[nodes_off[f]:nodes_off[f] + ns]
        Rf = 9 * ns + 1
        Fm = np.zeros((Rf, Rf))
        diag = np.zeros(9 * np_)
        if with_orig:
            hm = plan['hmap'][plan['f_hmap_off'][f]:plan['f_hmap_off'][f + 1]].reshape(ns, np_)
            for cs in range(np_):
                nc = nd[cs]
                blk = Hd[nc].copy()
                diag[9 * cs:9 * cs + 9] = np.diag(blk)
                blk[np.arange(9), np.arange(9)] = np.clip(np.diag(blk), lm_min, lm_max) * scale if clamp else 0.0
                Fm[9 * cs:9 * cs + 9, 9 * cs:9 * cs + 9] = blk
                Fm[Rf - 1, 9 * cs:9 * cs + 9] = -g[nc]
                for rs in range(cs + 1, ns):
                    h = hm[rs, cs]
                    if h >= 0:
                        b = Ho[h >> 1]
                        Fm[9 * rs:9 * rs + 9, 9 * cs:9 * cs + 9] = b.T if (h & 1) else b
        for k in range(plan['f_child_off'][f], plan['f_child_off'][f + 1]):
            c = int(plan['f_children'][k])
            if not child_filter(c):
                continue
            inv = plan['c_inv'][plan['c_inv_off'][k]:plan['c_inv_off'][k + 1]]
            nbc = int(plan['f_nb'][c])
            idx = np.full(9 * nbc + 1, -1)
            for s in range(ns):
                if inv[s] >= 0:
                    idx[9 * inv[s]:9 * inv[s] + 9] = np.arange(9 * s, 9 * s + 9)
            idx[9 * nbc] = Rf - 1
            Fm[np.ix_(idx, idx)] += np.tril(U[c])
        return np.tril(Fm), diag

    def eliminate(f, Fm):
        Cf = 9 * int(plan['f_np'][f])
        F11 = Fm[:Cf, :Cf]
        L11 = np.linalg.cholesky(F11 + np.tril(F11, -1).T)
        L21 = np.linalg.solve(L11, Fm[Cf:, :Cf].T).T
        Lp[f] = (L11, L21)
        U[f] = Fm[Cf:, Cf:] - np.tril(L21 @ L21.T)

    for f in order:                                         # private fronts of this rank
        if part[f] == rank:
            Fm, _ = front_matrix(f, True, True, lambda c: True)
            eliminate(f, Fm)
    shared = [int(f) for f in order if part[f] < 0]
    bases = {}
    for f in shared:                                        # partial panels + partial original diagonals
        Fm, diag = front_matrix(f, True, False, lambda c: part[c] == rank)
        bases[f] = (Fm, diag)
    buf = np.concatenate([np.concatenate([bases[f][0].ravel(), bases[f][1]]) for f in shared]) if shared else np.zeros(0)
    buf = allreduce(buf)
    off = 0
    for f in shared:                                        # redundant on every rank
        Rf = 9 * (int(plan['f_np'][f]) + int(plan['f_nb'][f])) + 1
        Cf = 9 * int(plan['f_np'][f])
        Fm = buf[off:off + Rf * Rf].reshape(Rf, Rf).copy()
        diag = buf[off + Rf * Rf:off + Rf * Rf + Cf]
        off += Rf * Rf + Cf
        Fm[np.arange(Cf), np.arange(Cf)] += np.clip(diag, lm_min, lm_max) * scale
        Fc, _ = front_matrix(f, False, False, lambda c: part[c] < 0)
        eliminate(f, Fm + Fc)
    D = np.zeros((N, 9))
    for f in order[::-1]:
        if not (part[f] < 0 or part[f] == rank):
            continue
        np_, nb = int(plan['f_np'][f]), int(plan['f_nb'][f])
        nd = nodes[nodes_off[f]:nodes_off[f] + np_ + nb]
        L11, L21 = Lp[f]
        t = L21[-1] - L21[:-1].T @ D[nd[np_:]].reshape(-1)
        D[nd[:np_]] = np.linalg.solve(L11.T, t).reshape(np_, 9)
    return D
