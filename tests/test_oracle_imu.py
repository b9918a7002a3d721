"""Pins oracle/imu_oracle.py: closed forms for constant rate / acceleration, a naive sequential integrator, golden window."""
import os

import numpy as np

from islam_b200 import synth
from oracle import imu_oracle, lie

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'imu_golden.npz'))


def test_constant_rate_and_acceleration_closed_form():
    F, h = 50, 0.01
    w = np.array([0.0, 0.0, 0.7]); a = np.array([0.3, 0.0, 0.0])
    gyro = np.tile(w, (F, 1)); acc = np.tile(a, (F, 1)); dt = np.full((F, 1), h)
    out = imu_oracle.preintegrate(dt, gyro, acc, np.zeros(3), np.array([0, 0, 0, 1.0]), np.zeros(3), gravity=0.0)
    # rotation: exact
    assert np.abs(lie.so3_log(out['rot'][-1]) - w * F * h).max() < 1e-12
    # gravity-free, zero rate => uniformly accelerated motion
    out0 = imu_oracle.preintegrate(dt, 0 * gyro, acc, np.ones(3), np.array([0, 0, 0, 1.0]), np.array([1.0, 0, 0]), 0.0)
    T = F * h
    assert np.abs(out0['vel'][-1] - (np.array([1.0, 0, 0]) + a * T)).max() < 1e-12
    assert np.abs(out0['pos'][-1] - (np.ones(3) + np.array([1.0, 0, 0]) * T + 0.5 * a * T * T)).max() < 1e-12


def test_gravity_cancels_for_a_static_level_imu():
    F, h, gv = 30, 0.01, 9.81007
    out = imu_oracle.preintegrate(np.full((F, 1), h), np.zeros((F, 3)), np.tile([0, 0, gv], (F, 1)), np.zeros(3),
                                  np.array([0, 0, 0, 1.0]), np.zeros(3), gv)
    assert np.abs(out['vel']).max() < 1e-12 and np.abs(out['pos']).max() < 1e-12


def test_integrate_matches_naive_sequential_integrator_and_golden():
    acc, gyr, dts, sync = GOLD['accels'].astype(np.float64), GOLD['gyros'].astype(np.float64), GOLD['dts'].astype(np.float64), GOLD['sync']
    init = dict(pos=GOLD['init_pos'], rot=GOLD['init_rot'], vel=GOLD['init_vel'])
    N = len(sync)
    p, r, c, v = imu_oracle.integrate(acc, gyr, dts, sync, 0, N - 1, init, float(GOLD['gravity']), False, np.float64)
    assert c == [] and p.shape == (N, 3) and r.shape == (N, 4)
    assert np.abs(p - GOLD['world_pos']).max() < 1e-12 and np.abs(v - GOLD['world_vel']).max() < 1e-12
    # naive integrator: same recurrences sample by sample in the world frame
    R, P, V = init['rot'].astype(np.float64), init['pos'].astype(np.float64), init['vel'].astype(np.float64)
    g = np.array([0, 0, float(GOLD['gravity'])])
    for f in range(N - 1):
        if sync[f] == sync[f + 1]:
            V = np.zeros(3)
        for k in range(sync[f], sync[f + 1]):
            Rn = lie.so3_mul(R, lie.so3_exp(gyr[k] * dts[k]))
            a = acc[k] - lie.so3_act(lie.so3_inv(Rn), g)
            wa = lie.so3_act(R, a)
            P = P + V * dts[k] + 0.5 * wa * dts[k] ** 2
            V = V + wa * dts[k]
            R = Rn
        assert np.abs(P - p[f + 1]).max() < 1e-9 and np.abs(V - v[f + 1]).max() < 1e-9
        assert np.abs(lie.quat_canon(R) - lie.quat_canon(r[f + 1])).max() < 1e-9
    pm, rm, _, vm = imu_oracle.integrate(acc, gyr, dts, sync, 0, N - 1, init, float(GOLD['gravity']), True, np.float64)
    assert pm.shape == (N - 1, 3)
    assert np.abs(pm - GOLD['motion_pos']).max() < 1e-12
    assert np.abs(pm[6]).max() == 0 and np.abs(vm[6]).max() == 0 and abs(abs(rm[6][3]) - 1) < 1e-6   # the gap frame (float32 init quaternion: |q| = 1 +- 6e-8)


def test_motion_mode_feeds_consistent_pvgo_deltas():
    """IMU deltas integrated from noise-free raw samples close the pvgo.py:42-51 residuals on the ground truth."""
    N = 30
    imu = synth.raw_imu(N, sig_a=0.0, sig_g=0.0)
    gt, gv, _, _ = synth.ground_truth(N)
    dp, dr, _, dv = imu_oracle.integrate(imu['accels'], imu['gyros'], imu['dts'], imu['rgb2imu_sync'], 0, N - 1,
                                         imu['init'], imu['gravity'], True, np.float64)
    rel = lie.so3_mul(lie.so3_inv(gt[:-1, 3:]), gt[1:, 3:])
    assert np.abs(lie.so3_log(lie.so3_mul(lie.so3_inv(dr), rel))).max() < 1e-6
    assert np.abs(dv - (gv[1:] - gv[:-1])).max() < 5e-3             # first-order integrator vs analytic trajectory
    assert np.abs(dp - (gt[1:, :3] - gt[:-1, :3] - gv[:-1] * 0.1)).max() < 5e-3
