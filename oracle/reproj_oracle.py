"""ORACLE (test infrastructure, not product code) — the optional 5th residual of the pose-velocity graph: the sparse
reprojection factor of /root/reference/pvgo.py:53-61 with /root/reference/dense_ba.py:276-305 (SparseReprojectionLoss).

    motion_i = X_i^-1 X_{i+1}                     pvgo.py:54-56
    motion_0 = 0.1  (all seven numbers)           pvgo.py:57   in-place overwrite: a constant, non-unit "pose"
    T_i = C^-1 motion_i C                         dense_ba.py:300   C = rgb2imu_pose
    err = point2pixel(point3d, K, T^-1) - target  dense_ba.py:302   (N points x 2 per pair)
    -> (M, 2N) rows ordered [u0, v0, u1, v1, ...] pvgo.py:59-60, information (loss_weight[4] / N)^2 I   pvgo.py:130-131,141-143

PINNED: tests/golden/reproj_golden.npz holds outputs of the reference class itself, run in the build container under a
small PyPose stand-in (tests/golden/make_reproj_golden.py); tests/test_oracle_reproj.py checks `residual` against them.
PyPose's LieTensor arithmetic is restated as in oracle/lie.py (quaternion formulas WITHOUT normalisation: the overwritten
motion_0 has |q| = 0.2 and goes through Mul / Inv / Act as is).  The Jacobian is the one PyPose's autograd yields: every op
on the path is a LieTensor op, so it is the true left-tangent derivative for pairs >= 1 and ZERO for pair 0 (index_put of a
constant); validated against finite differences in the same test file.

Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this module.
"""
import numpy as np

from . import lie


def _pair_points(nodes, rp):
    """Camera-frame points p' = T^-1 P of every pair (M, N, 3) and the world-frame points W = X_i C P (pairs >= 1)."""
    nodes = np.asarray(nodes)
    dt = nodes.dtype
    C = np.asarray(rp['rgb2imu'], dt)
    P = np.asarray(rp['point3d'], dt)                                        # (M, N, 3)
    motion = lie.se3_mul(lie.se3_inv(nodes[:-1]), nodes[1:])                 # pvgo.py:54-56
    motion[0] = 0.1                                                          # pvgo.py:57
    T = lie.se3_mul(lie.se3_mul(lie.se3_inv(C)[None], motion), C[None])      # dense_ba.py:300
    Ti = lie.se3_inv(T)
    pc = lie.se3_act(Ti[:, None, :], P)                                      # extrinsics.unsqueeze(-2) @ points
    return pc, Ti


def residual(nodes, rp):
    """(M, 2N) reprojection residuals."""
    fx, fy, cx, cy = [np.asarray(nodes).dtype.type(v) for v in rp['K']]
    pc, _ = _pair_points(nodes, rp)
    u = fx * pc[..., 0] / pc[..., 2] + cx
    v = fy * pc[..., 1] / pc[..., 2] + cy
    err = np.stack([u, v], -1) - np.asarray(rp['target'], pc.dtype)          # (M, N, 2)
    return err.reshape(err.shape[0], -1)


def jacobian(nodes, rp):
    """J_i (M, 2N, 6): d r / d(left tangent of X_i); d r / d(left tangent of X_{i+1}) = -J_i; pair 0: zero."""
    nodes = np.asarray(nodes)
    dt = nodes.dtype
    fx, fy, cx, cy = [dt.type(v) for v in rp['K']]
    C = np.asarray(rp['rgb2imu'], dt)
    P = np.asarray(rp['point3d'], dt)
    M, N = P.shape[:2]
    pc, _ = _pair_points(nodes, rp)
    Xi, Xj = nodes[:-1], nodes[1:]
    B = lie.se3_act(C[None, None, :], P)                                     # body-frame points C P
    W = lie.se3_act(Xi[:, None, :], B)                                       # world-frame points X_i C P
    RC = lie.so3_matrix(C[3:])
    Rj = lie.so3_matrix(Xj[:, 3:])
    A = np.einsum('ba,mcb->mac', RC, Rj)                                     # R_C^T R_j^T   (M,3,3)
    G = np.zeros((M, N, 3, 6), dt)                                           # dW / d delta_i = [I | -[W]x]
    G[..., 0, 0] = G[..., 1, 1] = G[..., 2, 2] = 1
    G[..., :, 3:] = -lie.skew(W)
    x, y, z = pc[..., 0], pc[..., 1], pc[..., 2]
    Pi = np.zeros((M, N, 2, 3), dt)
    Pi[..., 0, 0] = fx / z; Pi[..., 0, 2] = -fx * x / z ** 2
    Pi[..., 1, 1] = fy / z; Pi[..., 1, 2] = -fy * y / z ** 2
    J = np.einsum('mnab,mbc,mncd->mnad', Pi, A, G)                           # (M, N, 2, 6)
    J[0] = 0                                                                 # motion[0] is a constant (pvgo.py:57)
    return J.reshape(M, 2 * N, 6)
