"""ORACLE (test infrastructure, not product code) — CPU restatement of the IMU pre-integration path.

PARITY UNPINNED (PyPose absent, the reference ships no tests): restates pp.module.IMUPreintegrator.forward as
described in SURVEY.md Appendix A.5 and the per-frame loop of /root/reference/imu_integrator.py:69-164
(IMUModule.integrate), line by line, in NumPy.  Covariance propagation is skipped: the reference never reads it
(imu_integrator.py:84,88,164 — `covs` stays []).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.
"""
import numpy as np

from . import lie


def preintegrate(dt, gyro, acc, init_pos, init_rot, init_vel, gravity):
    """pp.module.IMUPreintegrator(...).forward(dt, gyro, acc, init_state) for one frame (A.5).

    dt (F,1) or (F,), gyro (F,3), acc (F,3).  Returns dict of (F,3)/(F,4)/(F,3) 'pos','rot','vel' for k=1..F."""
    dtype = acc.dtype
    dt = dt.reshape(-1, 1).astype(dtype)
    F = dt.shape[0]
    # dR_0 = I ; dR_{k+1} = dR_k * Exp(w_k dt_k)     (pp.cumprod, right multiplication)
    inc = lie.so3_exp((gyro * dt).astype(dtype))
    dR = np.zeros((F + 1, 4), dtype)
    dR[0] = [0, 0, 0, 1]
    for k in range(F):
        dR[k + 1] = lie.so3_mul(dR[k], inc[k])
    g = np.array([0, 0, gravity], dtype)
    # a_k = acc_k - (R0 dR_{k+1})^-1 g          (gravity seen through the END-of-step attitude)
    Rk1 = lie.so3_mul(init_rot[None].astype(dtype), dR[1:])
    a = acc - lie.so3_act(lie.so3_inv(Rk1), g[None])
    Ra = lie.so3_act(dR[:-1], a)                                   # dR_k a_k
    dv = np.zeros((F + 1, 3), dtype)
    dp = np.zeros((F + 1, 3), dtype)
    tt = np.zeros((F + 1, 1), dtype)
    for k in range(F):
        dp[k + 1] = dp[k] + dv[k] * dt[k] + dtype.type(0.5) * Ra[k] * dt[k] * dt[k]
        dv[k + 1] = dv[k] + Ra[k] * dt[k]
        tt[k + 1] = tt[k] + dt[k]
    R0 = init_rot[None].astype(dtype)
    rot = lie.so3_mul(R0, dR[1:])
    vel = init_vel[None] + lie.so3_act(R0, dv[1:])
    pos = init_pos[None] + lie.so3_act(R0, dp[1:]) + init_vel[None] * tt[1:]
    return {'pos': pos.astype(dtype), 'rot': rot.astype(dtype), 'vel': vel.astype(dtype)}


def integrate(accels, gyros, dts, rgb2imu_sync, st, end, init, gravity, motion_mode=False, dtype=np.float32):
    """IMUModule.integrate (imu_integrator.py:69-164).  init = dict(pos, rot, vel).
    Returns (poses (K,3), rots (K,4), covs [], vels (K,3)); world mode K = end-st+1, motion mode K = end-st."""
    dtype = np.dtype(dtype)
    accels = np.asarray(accels, dtype); gyros = np.asarray(gyros, dtype); dts = np.asarray(dts, dtype).reshape(-1, 1)
    if motion_mode:                                                   # prase_init: imu_integrator.py:14-18
        init_pos = np.zeros(3, dtype); init_vel = np.zeros(3, dtype)
    else:
        init_pos = np.asarray(init['pos'], dtype); init_vel = np.asarray(init['vel'], dtype)
    init_rot = np.asarray(init['rot'], dtype)
    if motion_mode:
        poses, rots, vels = [], [], []
    else:
        poses, rots, vels = [init_pos], [init_rot], [init_vel]        # imu_integrator.py:86-89
    state = {'pos': init_pos[None], 'rot': init_rot[None], 'vel': init_vel[None]}
    last = {'pos': init_pos, 'rot': init_rot, 'vel': init_vel}
    b0 = rgb2imu_sync[st]
    b1 = rgb2imu_sync[end] + 1
    d, gy, ac = dts[b0:b1], gyros[b0:b1], accels[b0:b1]
    for i in range(st, end):
        f0 = rgb2imu_sync[i] - b0
        f1 = rgb2imu_sync[i + 1] - b0
        if f0 == f1:                                                  # imu_integrator.py:134-140
            if motion_mode:
                state['pos'] = np.zeros((1, 3), dtype); state['vel'] = np.zeros((1, 3), dtype)
            else:
                state['vel'] = np.zeros((1, 3), dtype)
        else:
            state = preintegrate(d[f0:f1], gy[f0:f1], ac[f0:f1], last['pos'], last['rot'], last['vel'], gravity)
        poses.append(state['pos'][-1])
        vels.append(state['vel'][-1])
        if motion_mode:
            rots.append(lie.so3_mul(lie.so3_inv(last['rot']), state['rot'][-1]))
        else:
            rots.append(state['rot'][-1])
        last['rot'] = state['rot'][-1]
        if not motion_mode:
            last['pos'] = state['pos'][-1]
            last['vel'] = state['vel'][-1]
    return np.stack(poses), np.stack(rots), [], np.stack(vels)
