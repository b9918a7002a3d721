"""N>1 host logic on CPU: world_size-2 / -4 `gloo` process groups run the sharded multifrontal scheme (factor ownership,
partial shared panels, ONE all-reduce, redundant shared factorisation) in NumPy over the C++ plan and must reproduce the
dense solve — the same partition, maps and ownership rule the CUDA path uses (tests/test_gpu_pvgo.py covers the kernels)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mf_emul
from islam_b200 import synth
from oracle import pvgo_oracle as po


_GRAPHS = {'band8': lambda: synth.config2(N=260, band=8), 'lc': lambda: synth.config4(N=300, n_lc=5, min_gap=40),
           # enough closures for the DENSE root (csrc/dense_root.cuh), which all ranks then factor together
           'lcdense': lambda: synth.config4(N=420, n_lc=40, min_gap=40)}


def _partial_system(g, plan, rank):
    """H, g assembled from the factors `rank` owns only (ownership rule of csrc/pvgo.cu)."""
    lm = po.SparseLM(g, np.float64)
    res = lm._res()
    _, eo, pown = mf_emul.owners(plan, lm.edges)
    em, pm = eo == rank, pown == rank
    w = lm.w
    Jvo, Jrot = po.jacobian_blocks(lm.nodes, lm.vels, lm.edges, lm.poses, lm.drots, res[0], res[2])
    N = g.N
    H = np.zeros((9 * N, 9 * N))
    gg = np.zeros((N, 9))
    for e in np.where(em)[0]:
        i, j = lm.edges[e]
        S = w[0] * Jvo[e].T @ Jvo[e]
        q = w[0] * Jvo[e].T @ res[0][e]
        for (a, b, sg) in ((i, i, 1), (j, j, 1), (i, j, -1), (j, i, -1)):
            H[9 * a:9 * a + 6, 9 * b:9 * b + 6] += sg * S
        gg[j, :6] += q; gg[i, :6] -= q
    I3 = np.eye(3)
    for m in np.where(pm)[0]:
        a, b, dt = m, m + 1, lm.dts[m]
        J = np.zeros((9, 18))                                     # rows: adjvel, rot, transvel ; cols: node a (9), node b (9)
        J[0:3, 6:9] = I3; J[0:3, 15:18] = -I3
        J[3:6, 3:6] = -Jrot[m]; J[3:6, 12:15] = Jrot[m]
        J[6:9, 0:3] = -I3; J[6:9, 6:9] = -dt * I3; J[6:9, 9:12] = I3
        W = np.diag([w[1]] * 3 + [w[2]] * 3 + [w[3]] * 3)
        r = np.concatenate([res[1][m], res[2][m], res[3][m]])
        Hm, gm = J.T @ W @ J, J.T @ W @ r
        idx = np.concatenate([np.arange(9 * a, 9 * a + 9), np.arange(9 * b, 9 * b + 9)])
        H[np.ix_(idx, idx)] += Hm
        gg[a] += gm[:9]; gg[b] += gm[9:]
    return H, gg


def _worker(rank, world, port, name, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = _GRAPHS[name]()
    plan = mf_emul.get_plan(g.N, g.links, n_parts=world)
    H, gg = _partial_system(g, plan, rank)
    Hd, Ho = mf_emul.blocks_from_dense(H, plan, g.N)

    def allreduce(buf):
        t = torch.from_numpy(buf.copy())
        dist.all_reduce(t)
        return t.numpy()
    D = mf_emul.solve_sharded(plan, rank, Hd, Ho, gg, 1.0 + 1e-4, allreduce)
    var_part = plan['part'][plan['var_front']]               # every variable is reported by the rank that solves it
    mine = (var_part == rank) | ((var_part < 0) & (rank == 0))
    t = torch.from_numpy(np.where(mine[:, None], D.reshape(-1, 3), 0.0).reshape(-1, 9))
    dist.all_reduce(t)                                           # ShardedPVGO.get_state's gather
    Hs = torch.from_numpy(H.copy()); gs = torch.from_numpy(gg.copy())
    dist.all_reduce(Hs); dist.all_reduce(gs)
    if rank == 0:
        np.savez(os.path.join(out_dir, 'r.npz'), D=t.numpy(), H=Hs.numpy(), g=gs.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize('world,name', [(2, 'band8'), (4, 'band8'), (8, 'band8'), (2, 'lc'), (2, 'lcdense'), (4, 'lcdense')])
def test_sharded_scheme_reproduces_dense_solve(world, name, tmp_path):
    port = 29600 + world * 7 + len(name)
    mp.spawn(_worker, args=(world, port, name, str(tmp_path)), nprocs=world, join=True)
    r = np.load(tmp_path / 'r.npz')
    H, gg, D = r['H'], r['g'], r['D']
    # the per-rank partial systems sum to the oracle's full system
    g = _GRAPHS[name]()
    if name == 'lcdense':
        assert mf_emul.get_plan(g.N, g.links, n_parts=world)['dense_root'] >= 0
    lm = po.SparseLM(g, np.float64)
    Href, gref, _, _ = lm.assemble(lm._res())
    assert np.abs(H - Href.toarray()).max() < 1e-9 * np.abs(H).max()
    assert np.abs(gg - gref).max() < 1e-9 * max(1.0, np.abs(gref).max())
    scale = 1.0 + 1e-4
    A = H.copy()
    d = np.clip(np.diag(A), 1e-4, 1e32) * scale
    A[np.arange(len(d)), np.arange(len(d))] = d
    Dref = np.linalg.solve(A, -gg.reshape(-1)).reshape(-1, 9)
    assert np.abs(D - Dref).max() <= 1e-8 * np.abs(Dref).max()
