"""Drop-in behaviour at the reference's call surface: run_pvgo (pvgo.py:122-205) incl. autograd to the VO motions, and a
reference-style caller written against `import pypose as pp` (PoseVelGraph-shaped nn.Module + pp.optim.LM loop)."""
import numpy as np
import pytest
import torch

from islam_b200 import synth
from islam_b200.pvgo import run_pvgo
from oracle import lie, pvgo_oracle as po

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.as_tensor(a)


def test_run_pvgo_outputs_and_gradient():
    g = synth.window()
    vo = _t(g.vo_motions).cuda().requires_grad_(True)           # train.py: motions carry the VO net's graph, on cuda
    tl, rl, nodes, vels, covs = run_pvgo(_t(g.init_nodes), _t(g.init_vels), vo, _t(g.links), _t(g.dts), _t(g.imu_drots),
                                         _t(g.imu_dtrans), _t(g.imu_dvels), device='cuda:0', radius=g.radius,
                                         loss_weight=g.loss_weight, target='vo')
    assert nodes.device.type == 'cpu' and vels.device.type == 'cpu' and not nodes.requires_grad
    assert set(covs) == {'vo_rot', 'imu_rot', 'vo_trans', 'imu_vel', 'transvel'}
    assert covs['imu_rot'][0] == g.loss_weight[2] ** 2 and covs['vo_trans'].shape == (g.E,)
    ref = po.SparseLM(g, np.float64).run()
    rn, rv = ref.aligned(g.init_nodes[0])
    assert po.rel_pose_error(np.asarray(nodes), rn)['rel'] <= 1e-5
    rtl, rrl = ref.vo_loss()
    assert np.allclose(tl.detach().cpu().numpy(), rtl, rtol=2e-3, atol=1e-8)
    assert np.allclose(rl.detach().cpu().numpy(), rrl, rtol=2e-3, atol=1e-9)
    # train.py:280-283
    loss_bp = torch.cat((rl, tl))
    assert loss_bp.requires_grad
    loss_bp.backward(torch.ones_like(loss_bp))
    gt, gr = po.vo_loss_grad(ref.nodes, ref.edges, ref.poses)
    got = vo.grad.cpu().numpy()
    assert np.abs(got[:, 6]).max() == 0
    assert np.abs(got[:, :6] - (gt + gr)).max() < 5e-3 * max(1e-3, np.abs(gt + gr).max())


def test_run_pvgo_imu_target_and_errors():
    g = synth.window()
    tl, rl, nodes, vels, _ = run_pvgo(_t(g.init_nodes), _t(g.init_vels), _t(g.vo_motions), _t(g.links), _t(g.dts),
                                      _t(g.imu_drots), _t(g.imu_dtrans), _t(g.imu_dvels), loss_weight=g.loss_weight,
                                      target='imu')
    ref = po.SparseLM(g, np.float64).run()
    a, b = ref.imu_loss()
    assert np.allclose(tl.cpu().numpy(), a, rtol=5e-3, atol=1e-8) and np.allclose(rl.cpu().numpy(), b, rtol=5e-3, atol=1e-9)
    with pytest.raises(AttributeError):                             # pvgo.py:131 reads reproj.N
        run_pvgo(_t(g.init_nodes), _t(g.init_vels), _t(g.vo_motions), _t(g.links), _t(g.dts), _t(g.imu_drots),
                 _t(g.imu_dtrans), _t(g.imu_dvels), loss_weight=(1, 1, 1, 1, 1), reproj=object())
    with pytest.raises(Exception):                                  # len(dts) must be N-1 (pvgo.py:51 broadcast)
        run_pvgo(_t(g.init_nodes), _t(g.init_vels), _t(g.vo_motions), _t(g.links), _t(g.dts)[:-1], _t(g.imu_drots),
                 _t(g.imu_dtrans), _t(g.imu_dvels))


def test_reference_style_caller_on_the_shim():
    """Code shaped like /root/reference/pvgo.py:15-64,125-197, written against `import pypose as pp`."""
    import islam_b200.pypose_compat as ppc
    ppc.install()
    import pypose as pp
    import pypose.optim.solver as ppos
    import pypose.optim.strategy as ppost
    from pypose.optim.scheduler import StopOnPlateau

    class Graph(torch.nn.Module):
        def __init__(self, nodes, vels):
            super().__init__()
            self.nodes = pp.Parameter(nodes.clone())
            self.vels = torch.nn.Parameter(vels.clone())
            self.reproj = None

        def vo_loss(self, edges, poses):
            n1 = self.nodes[edges[:, 0]].detach()
            n2 = self.nodes[edges[:, 1]].detach()
            e = (poses.Inv() @ n1.Inv() @ n2).Log().tensor()
            return torch.sum(e[:, :3] ** 2, dim=1), torch.sum(e[:, 3:] ** 2, dim=1)

        def align_to(self, target, idx=0):
            source = self.nodes[idx].detach()
            vels = target.rotation() @ source.rotation().Inv() @ self.vels
            nodes = target @ source.Inv() @ self.nodes
            return nodes, vels

    g = synth.config1()
    dev = 'cuda:0'
    w = g.loss_weight
    mats = lambda c, k, n: torch.stack([torch.diag(torch.tensor([c] * k)) for _ in range(n)]).to(torch.float32).to(dev)
    weights = [mats(w[0] ** 2, 6, g.E), mats(w[1] ** 2, 3, g.M), mats(w[2] ** 2, 3, g.M), mats(w[3] ** 2, 3, g.M)]
    init_nodes = pp.SE3(_t(g.init_nodes))
    graph = Graph(init_nodes, _t(g.init_vels)).to(dev)
    assert isinstance(graph.nodes, pp.LieTensor) and graph.nodes.is_cuda
    opt = pp.optim.LM(graph, solver=ppos.Cholesky(), strategy=ppost.TrustRegion(radius=g.radius), min=1e-4, vectorize=True)
    sched = StopOnPlateau(opt, steps=10, patience=3, decreasing=1e-3, verbose=False)
    inp = (_t(g.links).to(dev), pp.SE3(_t(g.vo_motions)).to(dev), pp.SO3(_t(g.imu_drots)).to(dev), _t(g.imu_dtrans).to(dev),
           _t(g.imu_dvels).to(dev), _t(g.dts).unsqueeze(-1).to(dev))
    while sched.continual():
        loss = opt.step(input=inp, weight=weights)
        sched.step(loss)
    ref = po.SparseLM(g, np.float64).run()
    assert sched.steps == len(ref.history)
    assert abs(float(loss) - ref.history[-1]['loss']) < 1e-4 * ref.history[-1]['loss']
    vo = pp.SE3(_t(g.vo_motions).to(dev).requires_grad_(True))
    tl, rl = graph.vo_loss(inp[0], vo)
    rtl, rrl = ref.vo_loss()
    assert np.allclose(tl.detach().cpu().numpy(), rtl, rtol=5e-3, atol=1e-8)
    nodes, vels = graph.align_to(init_nodes[0].to(dev))
    rn, rv = ref.aligned(g.init_nodes[0])
    assert po.rel_pose_error(nodes.detach().cpu().numpy(), rn)['rel'] <= 1e-5
    assert np.abs(vels.detach().cpu().numpy() - rv).max() < 1e-4
    (tl.sum() + rl.sum()).backward()


def test_failed_cholesky_surfaces_as_info_1(capsys):
    """PyPose's "Linear solver failed. Breaking optimization step..." path: the parameters stay untouched,
    the step is abandoned, and the failure is visible to the caller (LM state info = 1), never silent."""
    from islam_b200.solver import PVGOSolver
    g = synth.window()
    s = PVGOSolver(g.N, g.links, device='cuda:0')
    s.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight)
    s.set_state(g.init_nodes, g.init_vels)
    s.lm_reset(radius=g.radius, lm_max=-1.0, max_steps=3, use_scheduler=0)      # every pivot diagonal clamped to -1: not SPD
    st = s.lm_run()
    assert st.info == 1 and st.steps_done == 3
    n, v = s.get_state()
    assert np.array_equal(n.cpu().numpy(), g.init_nodes) and np.array_equal(v.cpu().numpy(), g.init_vels)
    # through run_pvgo: a NaN measurement poisons J^T W J; the call returns (PyPose prints and breaks the step), info = 1
    bad = _t(g.vo_motions).clone()
    bad[3, 0] = float('nan')
    tl, rl, nodes, vels, _ = run_pvgo(_t(g.init_nodes), _t(g.init_vels), bad, _t(g.links), _t(g.dts), _t(g.imu_drots),
                                      _t(g.imu_dtrans), _t(g.imu_dvels), loss_weight=g.loss_weight)
    assert run_pvgo.last_state.info == 1
    assert 'Linear solver failed' in capsys.readouterr().out


def test_two_live_graphs_do_not_share_device_state():
    """Two PoseVelGraph objects over the same (N, links) — consecutive sliding windows — own separate solver handles."""
    from islam_b200.pvgo import PoseVelGraph
    g = synth.window()
    a = PoseVelGraph(_t(g.init_nodes), _t(g.init_vels), links=_t(g.links))
    shifted = g.init_nodes.copy()
    shifted[:, 0] += 5.0
    b = PoseVelGraph(_t(shifted), _t(g.init_vels), links=_t(g.links))
    assert a.solver is not b.solver
    assert np.array_equal(a.nodes.cpu().numpy(), g.init_nodes) and np.array_equal(b.nodes.cpu().numpy(), shifted)
    # forward() stages the measurements of THIS call: new dts / motions are not ignored
    args = (_t(g.links), _t(g.vo_motions), _t(g.imu_drots), _t(g.imu_dtrans), _t(g.imu_dvels), _t(g.dts))
    r1 = a(*args)
    r2 = a(args[0], args[1], args[2], args[3] + 0.25, args[4], args[5])
    assert np.allclose((r1[3] - r2[3]).cpu().numpy(), 0.25, atol=1e-6)
    with pytest.raises(Exception):
        a.vo_loss(_t(g.links)[:-1], _t(g.vo_motions)[:-1])
