"""Pins oracle/pvgo_oracle.py: closed-form Jacobian blocks vs finite differences and vs torch.autograd with PyPose's
left-tangent convention restated as custom autograd Functions; literal dense LM vs sparse twin; golden C1 run."""
import os

import numpy as np
import pytest
import torch

from islam_b200 import synth
from oracle import lie, pvgo_oracle as po

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'c1_golden.npz'))


def _graph_from_gold():
    g = synth.config1()
    for k in ('init_nodes', 'init_vels', 'vo_motions', 'links', 'dts', 'imu_drots', 'imu_dtrans', 'imu_dvels'):
        assert np.array_equal(getattr(g, k), GOLD[k]), f'generator drifted from the committed fixture: {k}'
    return g


def test_generator_is_pinned_and_sizes_match_survey():
    g = _graph_from_gold()
    assert (g.N, g.E, g.M, g.factors, g.rows) == (100, 102, 99, 300, 1503)
    g2 = synth.config2()
    assert (g2.N, g2.E, g2.M, g2.factors, g2.rows) == (5000, 39964, 4999, 49962, 284775)   # C2


def test_jacobian_blocks_vs_finite_differences():
    g = synth.config2(N=40, band=3)
    lm = po.SparseLM(g, np.float64)
    res = lm._res()
    Jvo, Jrot = po.jacobian_blocks(lm.nodes, lm.vels, lm.edges, lm.poses, lm.drots, res[0], res[2])
    eps = 1e-5
    f = lambda n: po.residuals(n, lm.vels, lm.edges, lm.poses, lm.drots, lm.dtrans, lm.dvels, lm.dts)
    for e in (0, 17, 60):
        i, j = lm.edges[e]
        for k in range(6):
            d = np.zeros(6); d[k] = eps
            np_, nm = lm.nodes.copy(), lm.nodes.copy()
            np_[j] = lie.se3_retract(np_[j], d); nm[j] = lie.se3_retract(nm[j], -d)
            fd = (f(np_)[0][e] - f(nm)[0][e]) / (2 * eps)
            assert np.abs(fd - Jvo[e][:, k]).max() < 1e-6
            np_, nm = lm.nodes.copy(), lm.nodes.copy()
            np_[i] = lie.se3_retract(np_[i], d); nm[i] = lie.se3_retract(nm[i], -d)
            fd = (f(np_)[0][e] - f(nm)[0][e]) / (2 * eps)
            assert np.abs(fd + Jvo[e][:, k]).max() < 1e-6                        # J_i = -J_j  (A.3)
    for m in (0, 11):
        for k in range(3):
            d = np.zeros(6); d[3 + k] = eps
            np_, nm = lm.nodes.copy(), lm.nodes.copy()
            np_[m + 1] = lie.se3_retract(np_[m + 1], d); nm[m + 1] = lie.se3_retract(nm[m + 1], -d)
            fd = (f(np_)[2][m] - f(nm)[2][m]) / (2 * eps)
            assert np.abs(fd - Jrot[m][:, k]).max() < 1e-6


def test_dense_jacobian_matches_autograd_with_pypose_convention():
    """torch.autograd over restated LieTensor ops whose backward returns left-tangent gradients (A.1) reproduces the
    dense J of the literal oracle, including the [I|0] translation quirk (A.3)."""
    g = synth.config2(N=12, band=2)
    dlm = po.DenseLM(g, np.float64)
    J = dlm.dense_J(dlm._res())
    N, E, M = dlm.N, dlm.E, dlm.M

    class LogSE3(torch.autograd.Function):
        @staticmethod
        def forward(ctx, X):
            x = torch.from_numpy(lie.se3_log(X.detach().numpy()))
            ctx.save_for_backward(x)
            return x

        @staticmethod
        def backward(ctx, g_):
            (x,) = ctx.saved_tensors
            Ji = torch.from_numpy(lie.se3_Jl_inv(x.numpy()))
            gx = torch.einsum('...k,...kj->...j', g_, Ji)
            return torch.cat([gx, torch.zeros_like(gx[..., :1])], -1)

    class MulSE3(torch.autograd.Function):
        @staticmethod
        def forward(ctx, A, B):
            ctx.save_for_backward(A)
            return torch.from_numpy(lie.se3_mul(A.detach().numpy(), B.detach().numpy()))

        @staticmethod
        def backward(ctx, g_):
            (A,) = ctx.saved_tensors
            Ad = torch.from_numpy(lie.se3_adj(A.detach().numpy()))
            gb = torch.einsum('...k,...kj->...j', g_[..., :6], Ad)
            return g_, torch.cat([gb, torch.zeros_like(gb[..., :1])], -1)

    class InvSE3(torch.autograd.Function):
        @staticmethod
        def forward(ctx, X):
            Y = torch.from_numpy(lie.se3_inv(X.detach().numpy()))
            ctx.save_for_backward(Y)
            return Y

        @staticmethod
        def backward(ctx, g_):
            (Y,) = ctx.saved_tensors
            Ad = torch.from_numpy(lie.se3_adj(Y.numpy()))
            gx = -torch.einsum('...k,...kj->...j', g_[..., :6], Ad)
            return torch.cat([gx, torch.zeros_like(gx[..., :1])], -1)

    nodes = torch.from_numpy(dlm.nodes).requires_grad_(True)
    vels = torch.from_numpy(dlm.vels).requires_grad_(True)
    edges = torch.from_numpy(dlm.edges)
    Z = torch.from_numpy(dlm.poses)

    def model(nodes, vels):
        n1, n2 = nodes[edges[:, 0]], nodes[edges[:, 1]]                 # plain indexing: raw gradient passthrough
        pg = LogSE3.apply(MulSE3.apply(MulSE3.apply(InvSE3.apply(Z), InvSE3.apply(n1)), n2))
        adj = torch.from_numpy(dlm.dvels) - torch.diff(vels, dim=0)
        tv = torch.diff(nodes[:, :3], dim=0) - (vels[:-1] * torch.from_numpy(dlm.dts)[:, None] + torch.from_numpy(dlm.dtrans))
        return pg, adj, tv

    out = model(nodes, vels)
    rows = [(out[0], 0), (out[1], 6 * E), (out[2], 6 * E + 6 * M)]
    for r, off in rows:
        flat = r.reshape(-1)
        for k in range(0, flat.numel(), max(1, flat.numel() // 25)):
            gn, gv = torch.autograd.grad(flat[k], (nodes, vels), retain_graph=True, allow_unused=True)
            row = np.concatenate([(gn if gn is not None else torch.zeros_like(nodes)).numpy().reshape(-1),
                                  (gv if gv is not None else torch.zeros_like(vels)).numpy().reshape(-1)])
            assert np.abs(row - J[off + k]).max() < 1e-6, (off, k)     # float32-rounded (non-unit, 1 +- 6e-8) quaternions: Ad(XY) ~ Ad(X)Ad(Y)


def test_dense_and_sparse_oracles_agree_and_match_golden():
    g = _graph_from_gold()
    d = po.DenseLM(g, np.float64).run(steps=5)
    s = po.SparseLM(g, np.float64).run(steps=5)
    hist = GOLD['history']
    for k, (a, b) in enumerate(zip(d.history, s.history)):
        assert a['rejects'] == b['rejects'] == int(hist[k, 2])
        assert abs(a['loss'] - b['loss']) < 1e-10 and abs(a['loss'] - hist[k, 0]) < 1e-10
        assert abs(a['damping'] - hist[k, 3]) < 1e-18
    assert po.rel_pose_error(s.nodes, d.nodes)['rel'] < 1e-12
    n, v = s.aligned(g.init_nodes[0].astype(np.float64))
    assert po.rel_pose_error(n, GOLD['nodes'])['rel'] < 1e-10
    assert np.abs(v - GOLD['vels']).max() < 1e-9
    tl, rl = s.vo_loss()
    assert np.allclose(tl, GOLD['trans_loss'], atol=1e-10) and np.allclose(rl, GOLD['rot_loss'], atol=1e-10)


def test_float32_oracle_stays_within_tolerance_on_c1():
    g = _graph_from_gold()
    s32 = po.SparseLM(g, np.float32).run(steps=5)
    n, _ = s32.aligned(g.init_nodes[0])
    assert po.rel_pose_error(n, GOLD['nodes'])['rel'] < 1e-5          # C1 can be held to 1e-5 in fp32


def test_rejection_path_and_scheduler():
    """A huge first damping radius forces rejected tries; StopOnPlateau stops on patience (A.4)."""
    g = synth.config2(N=60, band=2)
    lm = po.SparseLM(g, np.float64, radius=1e12)
    lm.run()
    assert 1 <= len(lm.history) <= 10
    assert all(h['loss'] <= h['last'] + 1e-9 or h['rejects'] >= 16 for h in lm.history)
    w = synth.window()
    lw = po.SparseLM(w, np.float64).run()
    assert lw.history[-1]['rejects'] == 16 and len(lw.history) == 2    # unweighted accept test burns 16 rejects


def test_vo_loss_gradient_formula():
    g = synth.config2(N=20, band=2)
    lm = po.SparseLM(g, np.float64).run(steps=2)
    P = lm.poses.copy()
    gt, gr = po.vo_loss_grad(lm.nodes, lm.edges, P)
    eps = 1e-6
    for e in (0, 5):
        for k in range(6):
            d = np.zeros(6); d[k] = eps
            Pp, Pm = P.copy(), P.copy()
            Pp[e] = lie.se3_retract(Pp[e], d); Pm[e] = lie.se3_retract(Pm[e], -d)
            tp, rp = lm.vo_loss(Pp); tm, rm = lm.vo_loss(Pm)
            assert abs((tp[e] - tm[e]) / (2 * eps) - gt[e, k]) < 1e-6
            assert abs((rp[e] - rm[e]) / (2 * eps) - gr[e, k]) < 1e-6
