"""Pins oracle/lie.py against independent SciPy implementations (tests/golden/lie_golden.npz) and identities."""
import os

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from oracle import lie

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'lie_golden.npz'))


def test_se3_exp_matches_expm():
    X = lie.se3_exp(G['xi'])
    ref = G['X'].copy()
    assert np.abs(X[:, :3] - ref[:, :3]).max() < 1e-9
    assert np.abs(lie.quat_canon(X[:, 3:]) - lie.quat_canon(ref[:, 3:])).max() < 1e-9


def test_se3_log_inverts_exp():
    xi = G['xi'][:40]
    back = lie.se3_log(lie.se3_exp(xi))
    # |phi| may exceed pi in the random draws: compare group elements instead of algebra coordinates
    assert np.abs(lie.se3_exp(back) - lie.se3_exp(xi)).max() < 1e-9 or \
        np.abs(lie.quat_canon(lie.se3_exp(back)[:, 3:]) - lie.quat_canon(lie.se3_exp(xi)[:, 3:])).max() < 1e-9
    small = G['xi'][40:48]
    assert np.abs(lie.se3_log(lie.se3_exp(small)) - small).max() < 1e-12


def test_adjoint_matches_matrix_form():
    assert np.abs(lie.se3_adj(G['X']) - G['Ad']).max() < 1e-9


def test_left_jacobian_inverse_matches_finite_differences_of_logm():
    n = int(G['n_jinv'])
    xi = lie.se3_log(G['X'][:n])
    J = lie.se3_Jl_inv(xi)
    assert np.abs(J - G['Jinv'][:n]).max() < 5e-6


def test_so3_against_scipy_rotation():
    rng = np.random.default_rng(0)
    phi = rng.standard_normal((200, 3))
    q = lie.so3_exp(phi)
    ref = Rotation.from_rotvec(phi).as_quat()
    assert np.abs(lie.quat_canon(q) - lie.quat_canon(ref)).max() < 1e-12
    assert np.abs(lie.so3_matrix(q) - Rotation.from_rotvec(phi).as_matrix()).max() < 1e-12
    p = rng.standard_normal((200, 3))
    assert np.abs(lie.so3_act(q, p) - Rotation.from_rotvec(phi).apply(p)).max() < 1e-12
    q2 = lie.so3_exp(rng.standard_normal((200, 3)))
    assert np.abs(lie.quat_canon(lie.so3_mul(q, q2)) -
                  lie.quat_canon((Rotation.from_quat(q) * Rotation.from_quat(q2)).as_quat())).max() < 1e-12


@pytest.mark.parametrize('theta', [0.0, 1e-12, 1e-7, 1e-4, 0.049, 0.051, 0.49, 0.51, 1.0, 3.0, np.pi - 1e-6])
def test_edge_angles(theta):
    phi = np.array([[0.3, -0.5, 0.81]])
    phi = phi / np.linalg.norm(phi) * theta
    q = lie.so3_exp(phi)
    assert abs(np.linalg.norm(q) - 1) < 1e-12
    assert np.abs(lie.so3_log(q) - phi).max() < 1e-9
    assert np.abs(lie.so3_log(-q) - phi).max() < 1e-9              # q and -q give the same phi (A.2)
    assert np.abs(lie.so3_Jl(phi) @ lie.so3_Jl_inv(phi) - np.eye(3)).max() < 1e-9
    for dt in (np.float32,):
        assert np.abs(lie.so3_log(lie.so3_exp(phi.astype(dt))).astype(np.float64) - phi).max() < 2e-6


def test_group_identities():
    rng = np.random.default_rng(1)
    X = lie.se3_exp(rng.standard_normal((50, 6)))
    Y = lie.se3_exp(rng.standard_normal((50, 6)))
    I = lie.se3_mul(X, lie.se3_inv(X))
    assert np.abs(I[:, :3]).max() < 1e-12 and np.abs(np.abs(I[:, 6]) - 1).max() < 1e-12
    p = rng.standard_normal((50, 3))
    assert np.abs(lie.se3_act(lie.se3_mul(X, Y), p) - lie.se3_act(X, lie.se3_act(Y, p))).max() < 1e-12
    # Ad(XY) = Ad(X) Ad(Y)
    assert np.abs(lie.se3_adj(lie.se3_mul(X, Y)) - lie.se3_adj(X) @ lie.se3_adj(Y)).max() < 1e-11
    # left retraction: Exp(d) X
    d = rng.standard_normal((50, 6)) * 0.1
    assert np.abs(lie.se3_retract(X, d) - lie.se3_mul(lie.se3_exp(d), X)).max() == 0
