"""GPU parity of the optional reprojection factor (/root/reference/pvgo.py:53-61,130-165 with
/root/reference/dense_ba.py:276-305) through the C ABI: residuals against outputs of the reference class itself
(tests/golden/reproj_golden.npz), normal equations and LM steps against the oracle, run_pvgo end to end, and the reference's
own pvgo.py + dense_ba.py executed unmodified over the shim with a reprojection loss object."""
import os
import types

import numpy as np
import pytest
import torch

from islam_b200 import synth
from islam_b200.pvgo import PoseVelGraph, run_pvgo
from islam_b200.solver import PVGOSolver
from oracle import pvgo_oracle as po, reproj_oracle as ro
from test_gpu_pvgo import _dense_from_blocks

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reproj_golden.npz'))
_t = torch.as_tensor


def _reproj_obj(rp, cls=types.SimpleNamespace):
    """An object shaped like dense_ba.SparseReprojectionLoss (attributes only)."""
    fx, fy, cx, cy = [float(v) for v in rp['K']]
    return cls(N=int(rp['point3d'].shape[1]), point3d=_t(rp['point3d']).cuda(), target=_t(rp['target']).cuda(),
               K=torch.tensor([fx, 0, cx, 0, fy, cy, 0, 0, 1], dtype=torch.float32).view(3, 3).cuda(),
               rgb2imu_pose=_t(rp['rgb2imu']).cuda())


def _graph_with_reproj(g, n_points, weight=2.0):
    rp = synth.reproj_data(g, n_points, weight=weight)
    g.extra['reproj'] = rp
    g.loss_weight = tuple(g.loss_weight[:4]) + (weight,)
    return g, rp


def test_residuals_match_reference_class_golden():
    for name in [str(c) for c in G['cases']]:
        rp = dict(point3d=G[f'{name}_point3d'], target=G[f'{name}_target'], K=G[f'{name}_K'], rgb2imu=G[f'{name}_rgb2imu'])
        nodes = G[f'{name}_nodes']
        N = nodes.shape[0]
        g = synth.config3(N=N) if N != 9 else synth.window()
        graph = PoseVelGraph(_t(nodes), _t(g.init_vels), _reproj_obj(rp), links=_t(g.links))
        graph._loss_weight = (1, 1, 1, 1, 1)
        out = graph(_t(g.links), _t(g.vo_motions), _t(g.imu_drots), _t(g.imu_dtrans), _t(g.imu_dvels), _t(g.dts))
        assert len(out) == 5 and out[4].shape == G[f'{name}_err'].shape            # pvgo.py:58-61
        want = G[f'{name}_err']
        assert np.abs(out[4].cpu().numpy() - want).max() <= 1e-4 * np.abs(want).max(), name
        # the mirror of the loss class, called like dense_ba.py:299-305 on LieTensors
        import islam_b200.pypose_compat as pp
        from islam_b200.dense_ba import SparseReprojectionLoss
        loss = SparseReprojectionLoss(_t(G[f'{name}_pts2d']), _t(G[f'{name}_depth']), _t(G[f'{name}_flow']),
                                      *[float(v) for v in rp['K']], pp.SE3(_t(rp['rgb2imu'])), device='cuda:0')
        assert np.abs(loss.point3d.cpu().numpy() - rp['point3d']).max() <= 1e-5 * np.abs(rp['point3d']).max()
        assert np.abs(loss.target.cpu().numpy() - rp['target']).max() <= 1e-5 * np.abs(rp['target']).max()
        nd = pp.SE3(_t(nodes).cuda())
        motion = nd[:-1].Inv() @ nd[1:]
        motion[0] = 0.1
        err = loss(motion)
        assert np.abs(err.reshape(N - 1, -1).cpu().numpy() - want).max() <= 1e-4 * np.abs(want).max()


@pytest.mark.parametrize('name,npts', [('win9', 24), ('band3_57', 40), ('C1', 7)])
def test_linearize_and_lm_steps_with_reprojection_match_oracle(name, npts):
    g = {'win9': synth.window, 'band3_57': lambda: synth.config2(N=57, band=3), 'C1': synth.config1}[name]()
    g, rp = _graph_with_reproj(g, npts)
    s = PVGOSolver(g.N, g.links)
    s.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight, reproj=_reproj_obj(rp))
    s.set_state(g.init_nodes, g.init_vels)
    s.linearize()
    res = [r.cpu().numpy() for r in s.residuals()]
    ref = po.SparseLM(g, np.float64)
    want = ref._res()
    assert len(res) == 5
    for a, b in zip(res, want):
        assert np.abs(a - b).max() <= 1e-4 * max(1.0, np.abs(b).max()), name
    Hd, Ho, gg, pairs = [t.cpu().numpy() for t in s.normal_equations()]
    Href, gref, _, _ = ref.assemble(want)
    Href = Href.toarray()
    H = _dense_from_blocks(g.N, Hd, Ho, pairs)
    assert np.abs(H - Href).max() <= 2e-4 * np.abs(Href).max(), (np.abs(H - Href).max(), np.abs(Href).max())
    assert np.abs(gg - gref).max() <= 1e-3 * max(1.0, np.abs(gref).max())
    steps = 3
    s.lm_reset(radius=g.radius, max_steps=steps, use_scheduler=0)
    for k in range(steps):
        ref.step()
        st = s.lm_step()
        h = ref.history[-1]
        # the constant residual of pair 0 (pvgo.py:57) dominates the loss: once a step improves it by less than float32 can
        # resolve in the sum of squares, accept / reject is rounding noise in ANY float32 implementation (the reference's too)
        if h['last'] - h['loss'] > 1e-5 * h['loss']:
            assert st.reject_count == h['rejects'], (k, st.reject_count, h)
        assert abs(st.loss - h['loss']) <= 1e-4 * max(1e-3, abs(h['loss']))
    n, _ = s.align(g.init_nodes[0])
    rn, _ = ref.aligned(g.init_nodes[0])
    assert po.rel_pose_error(n.cpu().numpy(), rn)['rel'] <= 1e-5
    # removing the factor again restores the four-group problem
    s.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight[:4])
    s.set_state(g.init_nodes, g.init_vels)
    s.linearize()
    assert len(s.residuals()) == 4


def test_run_pvgo_with_reprojection_and_reference_files():
    """run_pvgo(..., reproj=...) end to end: the repo's mirror, and the reference's pvgo.py + dense_ba.py unmodified over the shim."""
    g = synth.window()
    g, rp = _graph_with_reproj(g, 16)
    lm = po.SparseLM(g, np.float64).run()
    rn, rv = lm.aligned(g.init_nodes[0])
    args = lambda: (_t(g.init_nodes), _t(g.init_vels), _t(g.vo_motions), _t(g.links), _t(g.dts), _t(g.imu_drots),
                    _t(g.imu_dtrans), _t(g.imu_dvels))
    tl, rl, nodes, vels, covs = run_pvgo(*args(), device='cuda:0', radius=g.radius, loss_weight=g.loss_weight,
                                         reproj=_reproj_obj(rp))
    assert 'reproj' in covs and covs['reproj'][0] == (g.loss_weight[4] / 16) ** 2          # pvgo.py:131,204-205
    assert 1 <= run_pvgo.last_state.steps_done <= 10          # (the plateau test sees float32 noise of the constant pair-0 residual)
    assert po.rel_pose_error(np.asarray(nodes), rn)['rel'] <= 1e-5
    # the reference files
    from ref_loader import reference_modules
    with reference_modules() as mods:
        import pypose as pp
        robj = _reproj_obj(rp)
        robj.rgb2imu_pose = pp.SE3(robj.rgb2imu_pose)
        a = (pp.SE3(_t(g.init_nodes)), _t(g.init_vels), pp.SE3(_t(g.vo_motions).cuda()), _t(g.links), _t(g.dts),
             pp.SO3(_t(g.imu_drots)), _t(g.imu_dtrans), _t(g.imu_dvels))
        tl2, rl2, nodes2, vels2, covs2 = mods[0].run_pvgo(*a, device='cuda:0', radius=g.radius, loss_weight=g.loss_weight,
                                                          reproj=robj)
        assert po.rel_pose_error(nodes2.tensor().numpy(), rn)['rel'] <= 1e-5
        assert covs2['reproj'][0] == covs['reproj'][0]
        # the reference's own loss class, constructor and __call__, on the shim: its residual is what the fused kernel produces
        name = str(G['cases'][0])
        loss = mods[3].SparseReprojectionLoss(_t(G[f'{name}_pts2d']), _t(G[f'{name}_depth']).cuda(), _t(G[f'{name}_flow']).cuda(),   # (as TartanVO hands them over: on the GPU)
                                              *[float(v) for v in G[f'{name}_K']], pp.SE3(_t(G[f'{name}_rgb2imu'])), device='cuda:0')
        nd = pp.SE3(_t(G[f'{name}_nodes']).cuda())
        motion = nd[:-1].Inv() @ nd[1:]
        motion[0] = 0.1
        err = loss(motion).reshape(nd.shape[0] - 1, -1).cpu().numpy()
        assert np.abs(err - G[f'{name}_err']).max() <= 1e-4 * np.abs(G[f'{name}_err']).max()
