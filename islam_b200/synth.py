"""Synthetic pose-velocity graphs for BASELINE.json's configs.

Pure NumPy, float64 generation -> float32 cast.  The inputs it produces have exactly the shapes
`run_pvgo` receives at /root/reference/train.py:256-263:

    init_nodes (N,7) SE3 [t, qxyzw]   init_vels (N,3)   vo_motions (E,7)   links (E,2) int64
    dts (M,)   imu_drots (M,4)   imu_dtrans (M,3)   imu_dvels (M,3)       with M = N-1

This module contains its own tiny quaternion helpers (generation only; it is NOT the product's
Lie-group arithmetic, which lives in csrc/lie.cuh, and NOT the oracle).
"""
from dataclasses import dataclass, field
import numpy as np

LOSS_WEIGHT = (1.0, 0.1, 10.0, 0.1)       # /root/reference/run_kitti.sh:5


@dataclass
class PVGraph:
    name: str
    init_nodes: np.ndarray
    init_vels: np.ndarray
    vo_motions: np.ndarray
    links: np.ndarray
    dts: np.ndarray
    imu_drots: np.ndarray
    imu_dtrans: np.ndarray
    imu_dvels: np.ndarray
    gt_nodes: np.ndarray
    gt_vels: np.ndarray
    loss_weight: tuple = LOSS_WEIGHT
    radius: float = 1e4
    iters: int = 10
    extra: dict = field(default_factory=dict)

    @property
    def N(self):
        return self.init_nodes.shape[0]

    @property
    def E(self):
        return self.links.shape[0]

    @property
    def M(self):
        return self.N - 1

    @property
    def factors(self):          # F = E + 2M
        return self.E + 2 * self.M

    @property
    def rows(self):             # R = 6E + 9M
        return 6 * self.E + 9 * self.M


# ------------------------------------------------------------------ small float64 quaternion helpers
def _qmul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz], -1)


def _qinv(q):
    return q * np.array([-1.0, -1.0, -1.0, 1.0])


def _qrot(q, p):
    v, w = q[..., :3], q[..., 3:4]
    t = 2.0 * np.cross(v, p)
    return p + w * t + np.cross(v, t)


def _qexp(phi):
    th = np.linalg.norm(phi, axis=-1, keepdims=True)
    k = np.where(th < 1e-8, 0.5, np.sin(0.5 * th) / np.where(th < 1e-8, 1, th))
    return np.concatenate([k * phi, np.cos(0.5 * th)], -1)


def _q_from_R(R):
    """Batched rotation matrix -> quaternion xyzw (w >= 0 branch is enough here: angles < pi)."""
    from scipy.spatial.transform import Rotation
    return Rotation.from_matrix(R).as_quat()


def _se3_mul(A, B):
    return np.concatenate([A[..., :3] + _qrot(A[..., 3:], B[..., :3]), _qmul(A[..., 3:], B[..., 3:])], -1)


def _se3_inv(X):
    qi = _qinv(X[..., 3:])
    return np.concatenate([-_qrot(qi, X[..., :3]), qi], -1)


def _se3_exp_small(xi):
    """Exp for noise perturbations: t = Jl(phi) tau, first-order-accurate series is plenty."""
    tau, phi = xi[..., :3], xi[..., 3:]
    th2 = np.sum(phi * phi, -1, keepdims=True)
    c1 = 0.5 - th2 / 24.0
    c2 = 1.0 / 6.0 - th2 / 120.0
    kt = np.cross(phi, tau)
    t = tau + c1 * kt + c2 * np.cross(phi, kt)
    return np.concatenate([t, _qexp(phi)], -1)


# ------------------------------------------------------------------ ground-truth trajectory
def ground_truth(N, dt=0.1):
    """Figure-eight with vertical wobble, body x-axis along velocity, +-0.1 rad roll oscillation."""
    t = np.arange(N) * dt
    p = np.stack([20.0 * np.sin(0.1 * t), 10.0 * np.sin(0.2 * t), 0.5 * np.sin(0.3 * t)], -1)
    v = np.stack([2.0 * np.cos(0.1 * t), 2.0 * np.cos(0.2 * t), 0.15 * np.cos(0.3 * t)], -1)
    a = np.stack([-0.2 * np.sin(0.1 * t), -0.4 * np.sin(0.2 * t), -0.045 * np.sin(0.3 * t)], -1)
    x = v / np.linalg.norm(v, axis=-1, keepdims=True)
    up = np.array([0.0, 0.0, 1.0])
    y = np.cross(up, x)
    y /= np.linalg.norm(y, axis=-1, keepdims=True)
    z = np.cross(x, y)
    R = np.stack([x, y, z], -1)                       # columns = body axes in world
    q = _q_from_R(R)
    # make the quaternion track continuous (no sign flips) and add roll about body x
    for i in range(1, N):
        if np.dot(q[i], q[i - 1]) < 0:
            q[i] = -q[i]
    roll = 0.1 * np.sin(0.5 * t)
    qroll = _qexp(np.stack([roll, 0 * roll, 0 * roll], -1))
    q = _qmul(q, qroll)
    return np.concatenate([p, q], -1), v, a, t


def _finish(name, gt, gv, links, dt, rng, iters, sig_t=0.02, sig_r=0.002, sig_dv=0.002, sig_dp=0.002,
            imu=None):
    N = gt.shape[0]
    M = N - 1
    links = np.asarray(links, dtype=np.int64).reshape(-1, 2)
    E = links.shape[0]
    # VO / loop-closure measurements: Z = Xi^-1 Xj * Exp(noise)
    rel = _se3_mul(_se3_inv(gt[links[:, 0]]), gt[links[:, 1]])
    noise = rng.standard_normal((E, 6)) * np.array([sig_t] * 3 + [sig_r] * 3)
    Z = _se3_mul(rel, _se3_exp_small(noise))
    dts = np.full(M, dt)
    if imu is None:
        # perturb the exact pre-integrated quantities (C1/C2/C4)
        drot = _qmul(_qmul(_qinv(gt[:-1, 3:]), gt[1:, 3:]), _qexp(rng.standard_normal((M, 3)) * sig_r))
        dvel = (gv[1:] - gv[:-1]) + rng.standard_normal((M, 3)) * sig_dv
        dtrans = (gt[1:, :3] - gt[:-1, :3]) - gv[:-1] * dt + rng.standard_normal((M, 3)) * sig_dp
    else:
        drot, dtrans, dvel = imu
    # initial guess = dead-reckoned IMU chain (train.py:236-239)
    nodes = np.zeros((N, 7))
    vels = np.zeros((N, 3))
    nodes[0] = gt[0]
    vels[0] = gv[0]
    for i in range(M):
        nodes[i + 1, 3:] = _qmul(nodes[i, 3:], drot[i])
        nodes[i + 1, :3] = nodes[i, :3] + vels[i] * dts[i] + dtrans[i]
        vels[i + 1] = vels[i] + dvel[i]
        nodes[i + 1, 3:] /= np.linalg.norm(nodes[i + 1, 3:])
    f32 = np.float32
    return PVGraph(name, nodes.astype(f32), vels.astype(f32), Z.astype(f32), links, dts.astype(f32),
                   drot.astype(f32), dtrans.astype(f32), dvel.astype(f32), gt.astype(f32), gv.astype(f32),
                   iters=iters)


def chain_links(N, band=1):
    out = []
    for k in range(1, band + 1):
        i = np.arange(0, N - k)
        out.append(np.stack([i, i + k], -1))
    return np.concatenate(out, 0)


def config1(seed=0):
    """C1: 100 poses, chain 99 + 3 loop closures => E=102, M=99, F=300, R=1503; 5 fixed LM steps."""
    rng = np.random.default_rng(seed)
    gt, gv, _, _ = ground_truth(100)
    links = np.concatenate([chain_links(100, 1), np.array([[0, 50], [25, 75], [0, 99]])], 0)
    return _finish('C1', gt, gv, links, 0.1, rng, iters=5)


def config2(seed=0, N=5000, band=8):
    """C2: 5000 poses, VO band i->i+k, k=1..8 => E=39964, M=4999, F=49962; 10 fixed LM steps."""
    rng = np.random.default_rng(seed)
    gt, gv, _, _ = ground_truth(N)
    return _finish('C2' if N == 5000 and band == 8 else f'band{band}_N{N}', gt, gv, chain_links(N, band), 0.1,
                   rng, iters=10)


def raw_imu(N, per_frame=10, dt=0.1, gravity=9.81007, seed=0, sig_a=0.02, sig_g=0.002):
    """100 Hz raw IMU consistent with ground_truth(N): acc = R^T (p'' + g), gyro = body rate."""
    rng = np.random.default_rng(seed + 1)
    S = (N - 1) * per_frame
    h = dt / per_frame
    ts = np.arange(S + 1) * h
    p = np.stack([20.0 * np.sin(0.1 * ts), 10.0 * np.sin(0.2 * ts), 0.5 * np.sin(0.3 * ts)], -1)
    # orientation track at IMU rate
    v = np.stack([2.0 * np.cos(0.1 * ts), 2.0 * np.cos(0.2 * ts), 0.15 * np.cos(0.3 * ts)], -1)
    a = np.stack([-0.2 * np.sin(0.1 * ts), -0.4 * np.sin(0.2 * ts), -0.045 * np.sin(0.3 * ts)], -1)
    x = v / np.linalg.norm(v, axis=-1, keepdims=True)
    y = np.cross(np.array([0.0, 0.0, 1.0]), x)
    y /= np.linalg.norm(y, axis=-1, keepdims=True)
    z = np.cross(x, y)
    q = _q_from_R(np.stack([x, y, z], -1))
    for i in range(1, S + 1):
        if np.dot(q[i], q[i - 1]) < 0:
            q[i] = -q[i]
    roll = 0.1 * np.sin(0.5 * ts)
    q = _qmul(q, _qexp(np.stack([roll, 0 * roll, 0 * roll], -1)))
    # body rate from finite rotation between samples (exactly what Exp(w h) integrates back)
    dq = _qmul(_qinv(q[:-1]), q[1:])
    dq = np.where(dq[:, 3:4] < 0, -dq, dq)
    vn = np.linalg.norm(dq[:, :3], axis=-1, keepdims=True)
    ang = 2.0 * np.arctan2(vn, dq[:, 3:4])
    gyro = np.where(vn < 1e-12, 0.0, dq[:, :3] / np.where(vn < 1e-12, 1, vn) * ang) / h
    g = np.array([0.0, 0.0, gravity])
    acc = _qrot(_qinv(q[1:]), a[:-1] + g)            # gravity seen through the end-of-step attitude (A.5)
    gyro = gyro + rng.standard_normal(gyro.shape) * sig_g
    acc = acc + rng.standard_normal(acc.shape) * sig_a
    sync = np.arange(N) * per_frame                   # rgb2imu_sync[frame] = imu index
    return dict(accels=acc.astype(np.float32), gyros=gyro.astype(np.float32),
                dts=np.full(S, h, np.float32), rgb2imu_sync=sync, gravity=gravity,
                init=dict(pos=p[0].astype(np.float32), rot=q[0].astype(np.float32),
                          vel=v[0].astype(np.float32)))


def config3(seed=0, N=4541):
    """C3: KITTI-00 length chain (E = N-1); IMU deltas are filled by the caller from raw_imu()."""
    rng = np.random.default_rng(seed)
    gt, gv, _, _ = ground_truth(N)
    return _finish('C3', gt, gv, chain_links(N, 1), 0.1, rng, iters=10)


def config4(seed=0, N=50000, n_lc=2000, min_gap=100):
    """C4: chain N-1 + n_lc random loop closures with index gap > min_gap."""
    rng = np.random.default_rng(seed)
    gt, gv, _, _ = ground_truth(N)
    lc = []
    while len(lc) < n_lc:
        i, j = sorted(rng.integers(0, N, 2).tolist())
        if j - i > min_gap:
            lc.append((i, j))
    links = np.concatenate([chain_links(N, 1), np.array(lc)], 0)
    return _finish('C4' if N == 50000 else f'lc{n_lc}_N{N}', gt, gv, links, 0.1, rng, iters=10)


def window(seed=0, N=9):
    """C5-sized window: batch_size=8 => 9 poses, 8 VO edges, 8 IMU pairs (run_kitti.sh:8)."""
    rng = np.random.default_rng(seed)
    gt, gv, _, _ = ground_truth(N)
    return _finish(f'win{N}', gt, gv, chain_links(N, 1), 0.1, rng, iters=10)


def reproj_data(g, n_points=24, seed=0, sig_px=0.3, weight=2.0):
    """Inputs of the optional sparse reprojection factor (/root/reference/pvgo.py:53-61, dense_ba.py:276-305) for graph g:
    per consecutive pair, `n_points` 3-D points in the camera frame of pose i and their noisy pixel positions in the camera
    at pose i+1 (ground-truth motion), a camera-to-body transform `rgb2imu` and pin-hole intrinsics (fx, fy, cx, cy).
    Returns dict(point3d (M,N,3), target (M,N,2), K (4,), rgb2imu (7,), N, weight = loss_weight[4])."""
    rng = np.random.default_rng(seed + 77)
    M = g.N - 1
    fx, fy, cx, cy = 80.0, 82.0, 79.5, 55.5
    # camera looks along body x: camera z = body x, camera x = -body y, camera y = -body z (a usual rgb -> imu mounting) + offset
    Rc = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
    C = np.concatenate([[0.3, 0.05, -0.1], _q_from_R(Rc[None])[0]])
    z = 4.0 + 26.0 * rng.random((M, n_points))
    u = 160.0 * rng.random((M, n_points)); v = 112.0 * rng.random((M, n_points))
    P = np.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], -1)
    gt = g.gt_nodes.astype(np.float64)
    motion = _se3_mul(_se3_inv(gt[:-1]), gt[1:])
    T = _se3_mul(_se3_mul(_se3_inv(C)[None], motion), C[None])
    Ti = _se3_inv(T)
    Pc = _qrot(Ti[:, None, 3:], P) + Ti[:, None, :3]
    tgt = np.stack([fx * Pc[..., 0] / Pc[..., 2] + cx, fy * Pc[..., 1] / Pc[..., 2] + cy], -1)
    tgt = tgt + sig_px * rng.standard_normal(tgt.shape)
    return dict(point3d=P.astype(np.float32), target=tgt.astype(np.float32), K=np.array([fx, fy, cx, cy], np.float32),
                rgb2imu=C.astype(np.float32), N=n_points, weight=float(weight))
