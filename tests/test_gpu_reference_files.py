"""The reference's OWN back-end modules, byte for byte (tests/golden/ref_src/, staged from /root/reference by
tests/golden/vendor_reference.py), executed on top of `islam_b200.pypose_compat` installed as `pypose`:

    pvgo.py                    run_pvgo / PoseVelGraph / vo_loss / imu_loss / align_to      (pvgo.py:15-205)
    imu_integrator.py          IMUModule.integrate, both modes, with a gap frame              (imu_integrator.py:31-164)
    Datasets/transformation.py motion2pose_pypose / pose2motion_pypose / cvtSE3_pypose        (transformation.py:72-124)
    dense_ba.py                scale_from_disp_flow (+ its autograd into the motion), SparseReprojectionLoss (dense_ba.py:88-305)

Each result is compared with the repo's mirror (islam_b200.pvgo / .imu_integrator / .transformation) AND with the CPU oracle,
including the autograd of the outer losses into vo_motions (target='vo') and into imu_drots / imu_dvels (target='imu').
This is INTEGRATION.md section B run for real: the residual definitions, weights, optimiser configuration and call order come
from executing reference code, not from a restatement of it."""
import os

import numpy as np
import pytest
import torch

from islam_b200 import synth
from oracle import imu_oracle, lie, pvgo_oracle as po
from ref_loader import reference_modules

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, 'golden', 'imu_golden.npz'))
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def ref():
    """(pvgo, imu_integrator, transformation, dense_ba) modules of the reference, imported unmodified over the shim."""
    with reference_modules() as mods:
        yield mods


def _inputs(g, pp, grad=None):
    """The argument list of train.py:256-263: LieTensors for poses / motions / IMU rotations, plain tensors otherwise."""
    t = torch.as_tensor
    vo = t(g.vo_motions).to(DEV)
    dr, dv = t(g.imu_drots).clone(), t(g.imu_dvels).clone()
    if grad == 'vo':
        vo.requires_grad_(True)
    if grad == 'imu':
        dr.requires_grad_(True)
        dv.requires_grad_(True)
    return dict(init_nodes=pp.SE3(t(g.init_nodes)), init_vels=t(g.init_vels), vo_motions=pp.SE3(vo), links=t(g.links),
                dts=t(g.dts), imu_drots=pp.SO3(dr), imu_dtrans=t(g.imu_dtrans), imu_dvels=dv), vo, dr, dv


def _call(fn, a, g, target):
    return fn(a['init_nodes'], a['init_vels'], a['vo_motions'], a['links'], a['dts'], a['imu_drots'], a['imu_dtrans'],
              a['imu_dvels'], device=DEV, radius=g.radius, loss_weight=g.loss_weight, target=target)


@pytest.mark.parametrize('name', ['window9', 'C1'])
def test_reference_run_pvgo_vo_target(ref, name):
    import pypose as pp
    from islam_b200.pvgo import run_pvgo as ours
    g = synth.window() if name == 'window9' else synth.config1()
    a, vo, _, _ = _inputs(g, pp, 'vo')
    tl, rl, nodes, vels, covs = _call(ref[0].run_pvgo, a, g, 'vo')                 # reference file, unmodified
    b, vo2, _, _ = _inputs(g, pp, 'vo')
    tl2, rl2, nodes2, vels2, covs2 = _call(ours, b, g, 'vo')                       # the repo's mirror
    lm = po.SparseLM(g, np.float64).run()
    rn, rv = lm.aligned(g.init_nodes[0])
    assert isinstance(nodes, pp.LieTensor) and nodes.device.type == 'cpu' and not nodes.requires_grad
    for n_, v_ in ((nodes, vels), (nodes2, vels2)):
        assert po.rel_pose_error(n_.tensor().numpy() if hasattr(n_, 'tensor') else np.asarray(n_), rn)['rel'] <= 1e-5
        assert np.abs(np.asarray(v_) - rv).max() <= 1e-4 * max(1.0, np.abs(rv).max())
    assert np.abs(nodes.tensor().numpy() - np.asarray(nodes2.tensor() if hasattr(nodes2, 'tensor') else nodes2)).max() < 2e-6
    rtl, rrl = lm.vo_loss()
    for x, y in ((tl, rtl), (rl, rrl), (tl2, rtl), (rl2, rrl)):
        assert np.allclose(x.detach().cpu().numpy(), y, rtol=2e-3, atol=1e-8)
    assert set(covs) == set(covs2) and all(np.array_equal(covs[k], covs2[k]) for k in covs)
    # train.py:280-283: one-step gradient into the VO motions, through the reference's own vo_loss built from shim LieTensor ops
    for t_, r_, v_ in ((tl, rl, vo), (tl2, rl2, vo2)):
        loss_bp = torch.cat((r_, t_))
        assert loss_bp.requires_grad
        loss_bp.backward(torch.ones_like(loss_bp))
    gt, gr = po.vo_loss_grad(lm.nodes, lm.edges, lm.poses)
    want = gt + gr
    for v_ in (vo, vo2):
        got = v_.grad.cpu().numpy()
        assert np.abs(got[:, 6]).max() == 0
        assert np.abs(got[:, :6] - want).max() < 5e-3 * max(1e-3, np.abs(want).max())


def test_reference_run_pvgo_imu_target_backpropagates(ref):
    """pvgo.py:149-150,188-189: with target='imu' the loss is evaluated on the grad-carrying imu_drots / imu_dvels."""
    import pypose as pp
    from islam_b200.pvgo import run_pvgo as ours
    g = synth.window()
    a, _, dr, dv = _inputs(g, pp, 'imu')
    tl, rl, nodes, vels, _ = _call(ref[0].run_pvgo, a, g, 'imu')
    b, _, dr2, dv2 = _inputs(g, pp, 'imu')
    tl2, rl2, nodes2, vels2, _ = _call(ours, b, g, 'imu')
    lm = po.SparseLM(g, np.float64).run()
    ra, rb = lm.imu_loss()
    for x, y in ((tl, ra), (rl, rb), (tl2, ra), (rl2, rb)):
        assert np.allclose(x.detach().cpu().numpy(), y, rtol=5e-3, atol=1e-8)
    for t_, r_ in ((tl, rl), (tl2, rl2)):
        loss_bp = torch.cat((r_, t_))
        assert loss_bp.requires_grad                                       # train.py:282 would otherwise skip the backward
        loss_bp.backward(torch.ones_like(loss_bp))
    assert dr.grad is not None and dv.grad is not None and dr2.grad is not None and dv2.grad is not None
    # the mirror's fused backward (k_imu_loss) equals the reference's op-by-op LieTensor autograd
    sv, sr = max(1e-6, float(dv.grad.abs().max())), max(1e-6, float(dr.grad.abs().max()))
    assert float((dv.grad - dv2.grad).abs().max()) <= 2e-3 * sv
    assert float((dr.grad - dr2.grad).abs().max()) <= 2e-3 * sr
    assert float(dr.grad[:, 3].abs().max()) == 0 and float(dr2.grad[:, 3].abs().max()) == 0
    # ... and the analytic value: d|dv - diff v|^2 / d dv = 2 adjvelerr
    res = po.residuals(lm.nodes, lm.vels, lm.edges, lm.poses, lm.drots, lm.dtrans, lm.dvels, lm.dts)
    assert np.abs(dv.grad.numpy() - 2.0 * res[1]).max() <= 2e-3 * max(1e-6, np.abs(res[1]).max() * 2)


@pytest.mark.parametrize('motion', [False, True])
def test_reference_imu_module_with_gap(ref, motion):
    from islam_b200.imu_integrator import IMUModule as Ours
    init = dict(pos=GOLD['init_pos'], rot=GOLD['init_rot'], vel=GOLD['init_vel'])
    kw = dict(init=init, gravity=float(GOLD['gravity']), rgb2imu_sync=GOLD['sync'], device=DEV, denoise_accel=False,
              denoise_gyro=False)
    m_ref = ref[1].IMUModule(GOLD['accels'], GOLD['gyros'], GOLD['dts'], **kw)      # reference file, per-frame Python loop
    m_our = Ours(GOLD['accels'], GOLD['gyros'], GOLD['dts'], **kw)                  # fused kernels
    n = len(GOLD['sync'])
    p, r, c, v = m_ref.integrate(0, n - 1, init, motion_mode=motion)
    p2, r2, c2, v2 = m_our.integrate(0, n - 1, init, motion_mode=motion)
    rp, rr, _, rv = imu_oracle.integrate(GOLD['accels'], GOLD['gyros'], GOLD['dts'], GOLD['sync'], 0, n - 1, init,
                                         float(GOLD['gravity']), motion, np.float64)
    assert c == [] and c2 == [] and p.shape == p2.shape == rp.shape and tuple(r.shape) == tuple(r2.shape) == rr.shape
    qc = lambda q: lie.quat_canon(np.asarray(torch.as_tensor(q).numpy(), np.float64))
    for a_, b_ in ((p, p2), (v, v2)):
        assert np.abs(a_.numpy() - b_.numpy()).max() < 2e-5 * max(1.0, float(a_.abs().max()))
    assert np.abs(qc(r) - qc(r2)).max() < 1e-5
    assert np.abs(p.numpy() - rp).max() < 2e-5 * max(1.0, np.abs(rp).max())
    assert np.abs(v.numpy() - rv).max() < 2e-5 * max(1.0, np.abs(rv).max())
    assert np.abs(qc(r) - lie.quat_canon(rr)).max() < 1e-5


def test_reference_transformation_helpers(ref):
    import pypose as pp
    from islam_b200 import transformation as ours
    rng = np.random.default_rng(3)
    xi = rng.normal(size=(12, 6)) * np.array([0.5, 0.5, 0.5, 0.2, 0.2, 0.2])
    motion = torch.as_tensor(lie.se3_exp(xi), dtype=torch.float32, device=DEV)
    T0 = torch.as_tensor(lie.se3_exp(rng.normal(size=(1, 6)))[0], dtype=torch.float32)
    pose_ref = ref[2].motion2pose_pypose(pp.SE3(motion), pp.SE3(T0))            # Python loop of group products
    pose_our = ours.motion2pose_pypose(pp.SE3(motion), pp.SE3(T0))              # one prefix-product scan
    want = [T0.numpy().astype(np.float64)]
    for m in lie.se3_exp(xi):
        want.append(lie.se3_mul(want[-1][None], m[None])[0])
    want = np.stack(want)
    canon = lambda X: np.concatenate([X[:, :3], lie.quat_canon(X[:, 3:])], 1)
    for got in (pose_ref, pose_our):
        assert isinstance(got, pp.LieTensor) and got.shape == (13, 7)
        assert np.abs(canon(got.tensor().cpu().numpy().astype(np.float64)) - canon(want)).max() < 2e-5
    back_ref = ref[2].pose2motion_pypose(pose_ref)
    back_our = ours.pose2motion_pypose(pose_our)
    for got in (back_ref, back_our):
        assert np.abs(canon(got.tensor().cpu().numpy().astype(np.float64)) - canon(lie.se3_exp(xi))).max() < 2e-5
    six = torch.as_tensor(xi, dtype=torch.float32)
    a, b = ref[2].cvtSE3_pypose(six), ours.cvtSE3_pypose(six)
    assert np.abs(a.tensor().cpu().numpy() - b.tensor().cpu().numpy()).max() < 1e-6


def test_reference_window_loop_like_train_py(ref):
    """train.py:219-299 in miniature with the reference's modules end to end: IMU pre-integration (both modes) -> chained
    initial poses -> run_pvgo -> next window's init_state; compared window by window with the oracle."""
    import pypose as pp
    N, batch = 17, 8
    gt, gv, _, _ = synth.ground_truth(N)
    imu = synth.raw_imu(N)
    m = ref[1].IMUModule(imu['accels'], imu['gyros'], imu['dts'], init=imu['init'], gravity=imu['gravity'],
                         rgb2imu_sync=imu['rgb2imu_sync'], device=DEV, denoise_accel=False, denoise_gyro=False)
    rel = lie.se3_mul(lie.se3_inv(gt[:-1].astype(np.float64)), gt[1:].astype(np.float64))
    rng = np.random.default_rng(1)
    meas = lie.se3_mul(rel, lie.se3_exp(rng.normal(size=(N - 1, 6)) * np.array([.02, .02, .02, .002, .002, .002])))
    init_state = dict(pos=gt[0, :3].astype(np.float32), rot=gt[0, 3:].astype(np.float32), vel=gv[0].astype(np.float32))
    o_state = {k: v.copy() for k, v in init_state.items()}
    lw = (1.0, 0.1, 10.0, 0.1)
    for st in range(0, N - 1, batch):
        end = st + batch
        motions = pp.SE3(torch.as_tensor(meas[st:end], dtype=torch.float32, device=DEV))
        imu_trans, imu_rots, _, imu_vels = m.integrate(st, end, init_state, motion_mode=False)           # train.py:236
        imu_poses = pp.SE3(torch.cat((imu_trans, imu_rots.tensor()), axis=1))                            # :239
        imu_dtrans, imu_drots, _, imu_dvels = m.integrate(st, end, init_state, motion_mode=True)         # :244
        links = torch.stack([torch.arange(batch), torch.arange(1, batch + 1)], 1)
        dts = torch.full((batch,), 0.1)
        tl, rl, pgo_poses, pgo_vels, _ = ref[0].run_pvgo(imu_poses, imu_vels, motions, links, dts, imu_drots, imu_dtrans,
                                                        imu_dvels, device=DEV, radius=1e4, loss_weight=lw, target='vo')
        # the oracle's window: float64 pre-integration + SparseLM on the same measurements
        op, orr, _, ov = imu_oracle.integrate(imu['accels'], imu['gyros'], imu['dts'], imu['rgb2imu_sync'], st, end, o_state,
                                              imu['gravity'], False, np.float64)
        dp, dr, _, dv = imu_oracle.integrate(imu['accels'], imu['gyros'], imu['dts'], imu['rgb2imu_sync'], st, end, o_state,
                                             imu['gravity'], True, np.float64)
        og = synth.PVGraph(name='w', init_nodes=np.concatenate([op, orr], 1), init_vels=ov, vo_motions=meas[st:end],
                           links=links.numpy(), dts=dts.numpy().astype(np.float64), imu_drots=dr, imu_dtrans=dp, imu_dvels=dv,
                           gt_nodes=gt[st:end + 1], gt_vels=gv[st:end + 1], loss_weight=lw)
        lm = po.SparseLM(og, np.float64).run()
        rn, rv = lm.aligned(og.init_nodes[0])
        err = po.rel_pose_error(pgo_poses.tensor().numpy(), rn)
        assert err['rel'] <= 2e-5, (st, err)                    # float32 pre-integration feeds the float32 solve
        p = pgo_poses.tensor().numpy()
        init_state = dict(pos=p[-1][:3], rot=p[-1][3:] / np.linalg.norm(p[-1][3:]), vel=pgo_vels[-1].numpy())     # :297-299
        o_state = dict(pos=rn[-1][:3], rot=rn[-1][3:] / np.linalg.norm(rn[-1][3:]), vel=rv[-1])


def test_reference_scale_from_disp_flow_value_and_gradient(ref):
    """dense_ba.py:88-176 executed op by op on shim LieTensors (autograd in PyPose's convention) against the fused kernel,
    whose backward sums come out of the same pass: value, masks, and d(scale)/d(motion) as TartanVO.py:181 needs it."""
    import pypose as pp
    from islam_b200 import dense_ba as ours
    G = np.load(os.path.join(HERE, 'golden', 'scale_golden.npz'))
    for k in range(int(G['n'])):
        t = lambda a: torch.as_tensor(a).to(DEV)
        depth = t(G[f'{k}_depth']) if bool(G[f'{k}_has_depth']) else None
        mask = t(G[f'{k}_mask']) if bool(G[f'{k}_has_mask']) else None
        fx, fy, cx, cy = [float(x) for x in G[f'{k}_intr']]
        args = (t(G[f'{k}_disp']), t(G[f'{k}_flow']))
        m1 = t(G[f'{k}_motion']).requires_grad_(True)
        m2 = t(G[f'{k}_motion']).requires_grad_(True)
        s1, z1, k1, d1 = ref[3].scale_from_disp_flow(*args, pp.SE3(m1), fx, fy, cx, cy, float(G[f'{k}_baseline']), depth=depth,
                                                     mask=mask, disp_th=float(G[f'{k}_disp_th']))
        s2, z2, k2, d2 = ours.scale_from_disp_flow(*args, pp.SE3(m2), fx, fy, cx, cy, float(G[f'{k}_baseline']), depth=depth,
                                                   mask=mask, disp_th=float(G[f'{k}_disp_th']))
        assert torch.equal(k1, k2) and torch.equal(d1, d2)
        assert abs(float(s1) - float(s2)) <= 2e-4 * abs(float(s1))
        assert s1.requires_grad and s2.requires_grad
        s1.sum().backward()
        s2.sum().backward()
        g1, g2 = m1.grad.cpu().numpy(), m2.grad.cpu().numpy()
        assert g2[6] == 0 and abs(g1[6]) <= 1e-6 * max(1.0, np.abs(g1).max())
        assert np.abs(g1[:6] - g2[:6]).max() <= 5e-3 * max(1e-6, np.abs(g1[:6]).max()), (k, g1, g2)
