"""Generates the committed golden fixtures (run here, in the build container; PyPose is absent, so the vectors come
from independent SciPy implementations and from the float64 dense oracle — see oracle/ headers: PARITY UNPINNED).

    python tests/golden/make_golden.py

  lie_golden.npz   SO3/SE3 Exp/Log/Ad/Jl^-1 known answers from scipy.spatial.transform.Rotation + scipy.linalg.expm/logm
  c1_golden.npz    config C1 (100 poses / 300 factors) inputs, and per-step outputs of the literal dense LM oracle
  imu_golden.npz   a 40-frame raw IMU window with the sequential float64 integrator's world / motion outputs
"""
import os
import sys

import numpy as np
import scipy.linalg
from scipy.spatial.transform import Rotation

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from islam_b200 import synth                     # noqa: E402
from oracle import pvgo_oracle as po, imu_oracle  # noqa: E402


def hat6(xi):
    tau, phi = xi[:3], xi[3:]
    K = np.array([[0, -phi[2], phi[1]], [phi[2], 0, -phi[0]], [-phi[1], phi[0], 0]])
    M = np.zeros((4, 4)); M[:3, :3] = K; M[:3, 3] = tau
    return M


def lie_vectors():
    rng = np.random.default_rng(7)
    xi = np.concatenate([rng.standard_normal((40, 6)) * np.array([5, 5, 5, 1, 1, 1]),
                         rng.standard_normal((8, 6)) * 1e-5,
                         np.array([[1, 2, 3, 3.1, 0, 0], [1, 2, 3, 0, 0, 0], [0, 0, 0, 0.1, 0.2, 3.0]])])
    T = np.stack([scipy.linalg.expm(hat6(x)) for x in xi])
    q = Rotation.from_matrix(T[:, :3, :3]).as_quat()
    X = np.concatenate([T[:, :3, 3], q], 1)                     # SE3 = Exp(xi)
    Ad = np.zeros((len(xi), 6, 6))
    for i, t in enumerate(T):
        R, p = t[:3, :3], t[:3, 3]
        px = np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])
        Ad[i, :3, :3] = R; Ad[i, :3, 3:] = px @ R; Ad[i, 3:, 3:] = R
    # left Jacobian inverse by central differences of Log(Exp(d) Exp(xi)) in float64 via expm/logm
    def log6(Tm):
        Lm = np.real(scipy.linalg.logm(Tm))
        return np.array([Lm[0, 3], Lm[1, 3], Lm[2, 3], Lm[2, 1], Lm[0, 2], Lm[1, 0]])
    Jinv = np.zeros((len(xi), 6, 6))
    eps = 1e-6
    for i in range(40):           # generic inputs only (away from theta -> pi / 0 where logm is delicate)
        for k in range(6):
            d = np.zeros(6); d[k] = eps
            Jinv[i, :, k] = (log6(scipy.linalg.expm(hat6(d)) @ T[i]) - log6(scipy.linalg.expm(hat6(-d)) @ T[i])) / (2 * eps)
    return dict(xi=xi, X=X, Ad=Ad, Jinv=Jinv, n_jinv=40)


def c1_vectors():
    g = synth.config1()
    lm = po.DenseLM(g, np.float64).run(steps=5)
    n, v = lm.aligned(g.init_nodes[0].astype(np.float64))
    tl, rl = lm.vo_loss()
    hist = np.array([[h['loss'], h['last'], h['rejects'], h['damping']] for h in lm.history])
    return dict(init_nodes=g.init_nodes, init_vels=g.init_vels, vo_motions=g.vo_motions, links=g.links, dts=g.dts,
                imu_drots=g.imu_drots, imu_dtrans=g.imu_dtrans, imu_dvels=g.imu_dvels,
                loss_weight=np.array(g.loss_weight), radius=g.radius, history=hist, nodes=n, vels=v,
                trans_loss=tl, rot_loss=rl, nodes_raw=lm.nodes, vels_raw=lm.vels)


def imu_vectors():
    N = 41
    imu = synth.raw_imu(N, per_frame=10)
    sync = imu['rgb2imu_sync'].copy()
    sync[7] = sync[6]                  # frame 6 -> 7 has no IMU sample: exercises imu_integrator.py:134-140
    out = dict(accels=imu['accels'], gyros=imu['gyros'], dts=imu['dts'], sync=sync, gravity=imu['gravity'],
               init_pos=imu['init']['pos'], init_rot=imu['init']['rot'], init_vel=imu['init']['vel'])
    for mode, tag in ((False, 'world'), (True, 'motion')):
        p, r, _, v = imu_oracle.integrate(imu['accels'], imu['gyros'], imu['dts'], sync, 0, N - 1, imu['init'],
                                          imu['gravity'], motion_mode=mode, dtype=np.float64)
        out[f'{tag}_pos'], out[f'{tag}_rot'], out[f'{tag}_vel'] = p, r, v
    return out


if __name__ == '__main__':
    np.savez_compressed(os.path.join(HERE, 'lie_golden.npz'), **lie_vectors())
    np.savez_compressed(os.path.join(HERE, 'c1_golden.npz'), **c1_vectors())
    np.savez_compressed(os.path.join(HERE, 'imu_golden.npz'), **imu_vectors())
    print('golden fixtures written to', HERE)
