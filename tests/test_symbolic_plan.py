"""Host logic: the C++ symbolic analysis (ordering, elimination tree, gather maps) validated by a NumPy emulation of the
multifrontal numeric phase against a dense solve — no GPU needed."""
import numpy as np
import pytest

import mf_emul
from islam_b200 import synth
from oracle import pvgo_oracle as po

CASES = {
    'C1': lambda: synth.config1(),
    'band8': lambda: synth.config2(N=300, band=8),
    'band3_odd': lambda: synth.config2(N=57, band=3),
    'chain': lambda: synth.config3(N=131),
    'loop_closures': lambda: synth.config4(N=400, n_lc=6, min_gap=50),
    'dense_root': lambda: synth.config4(N=500, n_lc=40, min_gap=60),
    'window9': lambda: synth.window(),
    'two_nodes': lambda: synth.window(N=2),
}


@pytest.mark.parametrize('name', list(CASES))
@pytest.mark.parametrize('opts', [{}, {'leaf_max': 3, 'pivot_max': 2}, {'n_parts': 4}])
def test_plan_solves_like_dense(name, opts):
    g = CASES[name]()
    lm = po.SparseLM(g, np.float64)
    H, gg, _, _ = lm.assemble(lm._res())
    H = H.toarray()
    plan = mf_emul.get_plan(g.N, g.links, **opts)
    # every 3-dof variable (tau, phi, v of every pose) is a pivot exactly once; parents come later; levels respect the tree
    piv = np.concatenate([plan['vars'][plan['vars_off'][f]:plan['vars_off'][f] + plan['np'][f]] for f in range(plan['F'])])
    assert sorted(piv.tolist()) == list(range(3 * g.N))
    sparse = np.arange(plan['F']) != plan['dense_root']          # the dense root is not padded to whole 9-column steps
    assert np.all(plan['npad'][sparse] % 3 == 0) and np.all(plan['npad'] >= plan['np']) and np.all(plan['npad'] - plan['np'] < 3)
    par = plan['parent']
    assert all(p == -1 or p > f for f, p in enumerate(par))
    assert all(p == -1 or plan['level'][p] > plan['level'][f] for f, p in enumerate(par))
    Hd, Ho = mf_emul.blocks_from_dense(H, plan, g.N)
    scale = 1.0 + 1e-4
    D = mf_emul.solve(plan, Hd, Ho, gg, scale)
    A = H.copy()
    d = np.clip(np.diag(A), 1e-4, 1e32) * scale
    A[np.arange(len(d)), np.arange(len(d))] = d
    Dref = np.linalg.solve(A, -gg.reshape(-1)).reshape(-1, 9)
    assert np.abs(D - Dref).max() <= 1e-9 * np.abs(Dref).max()


def test_separators_are_trimmed_to_one_velocity():
    """The point of the variable-level ordering: a separator of the band-b chain holds tau/phi of b poses but ONE velocity."""
    g = synth.config2(N=600, band=8)
    plan = mf_emul.get_plan(g.N, g.links)
    top = int(np.argmax(plan['level']))                         # the root separator
    vs = plan['vars'][plan['vars_off'][top]:plan['vars_off'][top] + plan['np'][top]]
    assert len(vs) == 2 * 8 + 1 and np.sum(vs % 3 == 2) == 1
    assert len(np.unique(vs[vs % 3 != 2] // 3)) == 8
    assert plan['max_cols'] <= 72 and plan['n_levels'] <= 8


def test_multi_window_partition_is_consistent():
    g = synth.config2(N=600, band=8)
    plan = mf_emul.get_plan(g.N, g.links, n_parts=4)
    part = plan['part']
    assert set(part.tolist()) == {-1, 0, 1, 2, 3}
    # private fronts only see variables of their own window or shared variables
    var_part = part[plan['var_front']]
    for f in range(len(part)):
        if part[f] < 0:
            continue
        vs = plan['vars'][plan['vars_off'][f]:plan['vars_off'][f + 1]]
        assert set(var_part[vs[vs >= 0]].tolist()) <= {-1, int(part[f])}
    # shared fronts' ancestors are shared
    for f, p in enumerate(plan['parent']):
        if part[f] < 0 and p >= 0:
            assert part[p] < 0
    # windows are contiguous index ranges
    for w in range(4):
        idx = np.where(var_part == w)[0]
        assert set(var_part[idx.min():idx.max() + 1].tolist()) <= {w, -1}
    # every factor touches private variables of at most one window (ownership rule of islam_pvgo_create)
    vp, eo, po_ = mf_emul.owners(plan, g.links)
    for (i, j), o in zip(g.links, eo):
        assert set(np.concatenate([vp[i, :2], vp[j, :2]]).tolist()) <= {-1, int(o)}
    for i, o in enumerate(po_):
        assert set(np.concatenate([vp[i], vp[i + 1]]).tolist()) <= {-1, int(o)}


def test_invalid_graphs_are_rejected():
    import ctypes as C
    from islam_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    bad = np.array([[0, 5]], dtype=np.int64)
    assert L.islam_plan_build(C.byref(h), 3, 1, bad.ctypes.data, None) < 0          # endpoint out of range
    self_loop = np.array([[1, 1]], dtype=np.int64)
    assert L.islam_plan_build(C.byref(h), 3, 1, self_loop.ctypes.data, None) < 0
    assert L.islam_plan_build(C.byref(h), 1, 0, None, None) < 0                      # fewer than two poses


def _random_graph(seed):
    """Random edge lists: sparse or dense bands, missing chain edges, duplicated and reversed edges, a few long edges."""
    rng = np.random.default_rng(seed)
    N = int(rng.integers(2, 70))
    links = []
    band = int(rng.integers(1, 6))
    for k in range(1, band + 1):
        for i in range(N - k):
            if rng.random() < 0.7:
                links.append((i, i + k) if rng.random() < 0.8 else (i + k, i))
    for _ in range(int(rng.integers(0, 4))):                      # loop closures beyond band_max (default 16)
        a, b = sorted(rng.integers(0, N, 2).tolist())
        if b - a > 16:
            links.append((a, b))
    links += links[:int(rng.integers(0, 3))]                       # duplicates
    links = np.array(links, dtype=np.int64).reshape(-1, 2)
    gt, gv, _, _ = synth.ground_truth(N)
    return synth._finish(f'rand{seed}', gt, gv, links, 0.1, rng, 3)


@pytest.mark.parametrize('seed', range(24))
def test_random_graphs_plan_solves_like_dense(seed):
    """Property test of the ordering / trimming / push maps on arbitrary structures (not only the band patterns above)."""
    g = _random_graph(seed)
    lm = po.SparseLM(g, np.float64, solver='splu')
    H, gg, _, _ = lm.assemble(lm._res())
    H = H.toarray()
    opts = [{}, {'leaf_max': 2, 'pivot_max': 1}, {'n_parts': 2}, {'band_max': 2}][seed % 4]
    plan = mf_emul.get_plan(g.N, g.links, **opts)
    piv = np.concatenate([plan['vars'][plan['vars_off'][f]:plan['vars_off'][f] + plan['np'][f]] for f in range(plan['F'])])
    assert sorted(piv.tolist()) == list(range(3 * g.N))
    Hd, Ho = mf_emul.blocks_from_dense(H, plan, g.N)
    scale = 1.0 + 1e-4
    D = mf_emul.solve(plan, Hd, Ho, gg, scale)
    A = H.copy()
    d = np.clip(np.diag(A), 1e-4, 1e32) * scale
    A[np.arange(len(d)), np.arange(len(d))] = d
    Dref = np.linalg.solve(A, -gg.reshape(-1)).reshape(-1, 9)
    assert np.abs(D - Dref).max() <= 1e-8 * max(np.abs(Dref).max(), 1e-12)
