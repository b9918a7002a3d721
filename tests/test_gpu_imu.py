"""GPU parity of the fused IMU pre-integration (csrc/imu.cu via IMUModule.integrate) against the oracle's per-frame loop."""
import os

import numpy as np
import pytest
import torch

from islam_b200 import synth
from islam_b200.imu_integrator import IMUModule
from oracle import imu_oracle, lie

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'imu_golden.npz'))


def _module(imu, sync=None):
    return IMUModule(imu['accels'], imu['gyros'], imu['dts'], init=imu['init'], gravity=imu['gravity'],
                     rgb2imu_sync=imu['rgb2imu_sync'] if sync is None else sync, device='cuda:0',
                     denoise_accel=False, denoise_gyro=False)


def _quat_close(a, b, tol):
    return np.abs(lie.quat_canon(np.asarray(a, np.float64)) - lie.quat_canon(np.asarray(b, np.float64))).max() < tol


@pytest.mark.parametrize('motion', [False, True])
def test_golden_window_with_gap(motion):
    imu = dict(accels=GOLD['accels'], gyros=GOLD['gyros'], dts=GOLD['dts'], gravity=float(GOLD['gravity']),
               init=dict(pos=GOLD['init_pos'], rot=GOLD['init_rot'], vel=GOLD['init_vel']))
    sync = GOLD['sync']
    m = _module(imu, sync)
    N = len(sync)
    p, r, c, v = m.integrate(0, N - 1, imu['init'], motion_mode=motion)
    tag = 'motion' if motion else 'world'
    assert c == [] and p.shape == GOLD[f'{tag}_pos'].shape and tuple(r.shape) == GOLD[f'{tag}_rot'].shape
    assert p.device.type == 'cpu' and v.device.type == 'cpu'
    assert np.abs(p.numpy() - GOLD[f'{tag}_pos']).max() < 2e-5 * max(1.0, np.abs(GOLD[f'{tag}_pos']).max())
    assert np.abs(v.numpy() - GOLD[f'{tag}_vel']).max() < 2e-5 * max(1.0, np.abs(GOLD[f'{tag}_vel']).max())
    assert _quat_close(torch.as_tensor(r).numpy(), GOLD[f'{tag}_rot'], 1e-5)      # float32 chain of ~400 products


def test_kitti_length_trajectory_both_modes():
    """C3's IMU leg: 4541 frames, 45 400 samples at 100 Hz vs the float64 per-frame loop."""
    N = 4541
    imu = synth.raw_imu(N)
    m = _module(imu)
    for motion in (False, True):
        p, r, _, v = m.integrate(0, N - 1, imu['init'], motion_mode=motion)
        rp, rr, _, rv = imu_oracle.integrate(imu['accels'], imu['gyros'], imu['dts'], imu['rgb2imu_sync'], 0, N - 1,
                                             imu['init'], imu['gravity'], motion, np.float64)
        scale = max(1.0, np.abs(rp).max())
        assert np.abs(p.numpy() - rp).max() < 1e-4 * scale, (motion, np.abs(p.numpy() - rp).max(), scale)
        assert np.abs(v.numpy() - rv).max() < 1e-4 * max(1.0, np.abs(rv).max())
        assert _quat_close(torch.as_tensor(r).numpy(), rr, 2e-5)


def test_sub_window_and_single_frame():
    imu = synth.raw_imu(60)
    m = _module(imu)
    init = dict(pos=np.array([1., 2, 3], np.float32), rot=np.array([0, 0, 0.6, 0.8], np.float32), vel=np.array([.1, .2, .3], np.float32))
    for st, end in ((10, 18), (5, 6), (0, 59)):
        for motion in (False, True):
            p, r, _, v = m.integrate(st, end, init, motion_mode=motion)
            rp, rr, _, rv = imu_oracle.integrate(imu['accels'], imu['gyros'], imu['dts'], imu['rgb2imu_sync'], st, end, init,
                                                 imu['gravity'], motion, np.float64)
            assert p.shape == rp.shape
            assert np.abs(p.numpy() - rp).max() < 1e-5 * max(1.0, np.abs(rp).max())
            assert np.abs(v.numpy() - rv).max() < 1e-5 * max(1.0, np.abs(rv).max())


def test_bias_subtraction_matches_reference_flags():
    imu = synth.raw_imu(20)
    ab, gb = np.array([0.01, -0.02, 0.03], np.float32), np.array([0.001, 0.002, -0.001], np.float32)
    m = IMUModule(imu['accels'], imu['gyros'], imu['dts'], accel_bias=ab, gyro_bias=gb, init=imu['init'],
                  gravity=imu['gravity'], rgb2imu_sync=imu['rgb2imu_sync'], device='cuda:0')   # denoise_* default True
    p, r, _, v = m.integrate(0, 19, imu['init'], motion_mode=True)
    rp, rr, _, rv = imu_oracle.integrate(imu['accels'] - ab, imu['gyros'] - gb, imu['dts'], imu['rgb2imu_sync'], 0, 19,
                                         imu['init'], imu['gravity'], True, np.float64)
    assert np.abs(p.numpy() - rp).max() < 1e-5 and np.abs(v.numpy() - rv).max() < 1e-5


@pytest.mark.parametrize('F', [1, 10, 37])
def test_pp_module_imupreintegrator_forward_every_k(F):
    """pp.module.IMUPreintegrator.forward as /root/reference/imu_integrator.py:55-56,146 calls it: the state
    after EVERY sample k = 1..F (offsets = arange), with an explicit init_state and with the constructor's state."""
    import islam_b200.pypose_compat as pp
    rng = np.random.default_rng(F)
    dt = (0.01 + 0.002 * rng.random((F, 1))).astype(np.float32)
    gyro = (0.3 * rng.normal(size=(F, 3))).astype(np.float32)
    acc = (np.array([0.2, -0.1, 9.8]) + 0.5 * rng.normal(size=(F, 3))).astype(np.float32)
    pos, vel = np.array([1.0, -2.0, 0.5], np.float32), np.array([0.7, 0.1, -0.2], np.float32)
    rot = lie.so3_exp(np.array([[0.3, -0.2, 0.9]]))[0].astype(np.float32)
    t = lambda a: torch.as_tensor(a).cuda()
    integ = pp.module.IMUPreintegrator(t(pos), pp.SO3(t(rot)), t(vel), gravity=9.81007).to('cuda:0')
    want = imu_oracle.preintegrate(dt.astype(np.float64), gyro.astype(np.float64), acc.astype(np.float64),
                                   pos.astype(np.float64), rot.astype(np.float64), vel.astype(np.float64), 9.81007)
    for init_state in ({'pos': t(pos), 'rot': pp.SO3(t(rot)), 'vel': t(vel)}, None):
        out = integ(dt=t(dt), gyro=t(gyro), acc=t(acc), init_state=init_state)
        assert out['pos'].shape == (1, F, 3) and out['rot'].shape == (1, F, 4) and out['vel'].shape == (1, F, 3)
        assert isinstance(out['rot'], pp.LieTensor) and out['rot'].ltype is pp.SO3_type and out['pos'].is_cuda
        assert np.abs(out['pos'][0].cpu().numpy() - want['pos']).max() < 1e-5 * max(1.0, np.abs(want['pos']).max())
        assert np.abs(out['vel'][0].cpu().numpy() - want['vel']).max() < 1e-5 * max(1.0, np.abs(want['vel']).max())
        assert _quat_close(out['rot'][0].cpu().numpy(), want['rot'], 1e-5)
        # the last sample is what IMUModule.integrate keeps (imu_integrator.py:148-153)
        assert out['pos'][..., -1, :].squeeze().shape == (3,)
