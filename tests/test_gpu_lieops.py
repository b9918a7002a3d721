"""GPU parity of the elementwise LieTensor kernels (csrc/lieops.cu through the shim) against oracle/lie.py, forward and
left-tangent backward."""
import numpy as np
import pytest
import torch

import islam_b200.pypose_compat as pp
from oracle import lie

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _rand_se3(n, seed=0, scale=1.0):
    rng = np.random.default_rng(seed)
    return lie.se3_exp(rng.standard_normal((n, 6)) * np.array([3, 3, 3, scale, scale, scale])).astype(np.float32)


@pytest.mark.parametrize('dev', ['cuda:0', 'cpu'])
def test_forward_maps(dev):
    X = _rand_se3(257, 1)
    Y = _rand_se3(257, 2)
    x, y = pp.SE3(torch.tensor(X, device=dev)), pp.SE3(torch.tensor(Y, device=dev))
    f = lambda t: t.tensor().cpu().numpy().astype(np.float64) if isinstance(t, pp.LieTensor) else t.cpu().numpy().astype(np.float64)
    Xd, Yd = X.astype(np.float64), Y.astype(np.float64)
    assert np.abs(f(x.Inv()) - lie.se3_inv(Xd)).max() < 2e-5
    assert np.abs(f(x @ y) - lie.se3_mul(Xd, Yd)).max() < 2e-5
    assert np.abs(f(x.Log()) - lie.se3_log(Xd)).max() < 2e-5
    xi = lie.se3_log(Xd).astype(np.float32)
    assert np.abs(f(pp.se3(torch.tensor(xi, device=dev)).Exp()) - lie.se3_exp(xi.astype(np.float64))).max() < 2e-5
    p = np.random.default_rng(3).standard_normal((257, 3)).astype(np.float32)
    assert np.abs(f(x @ torch.tensor(p, device=dev)) - lie.se3_act(Xd, p.astype(np.float64))).max() < 2e-5
    r = x.rotation()
    assert r.ltype is pp.SO3_type
    assert np.abs(f(r.Inv() @ r)[:, :3]).max() < 1e-6
    assert np.abs(f(r.Log()) - lie.so3_log(Xd[:, 3:])).max() < 2e-6
    assert (x @ y).device.type == torch.device(dev).type


def test_small_and_large_angles():
    phi = np.array([[0, 0, 0], [1e-9, 0, 0], [1e-4, 2e-4, 0], [0.4, 0.1, 0.2], [0, 3.1, 0], [2.0, 2.0, 1.0]], np.float32)
    q = pp.so3(torch.tensor(phi, device=DEV)).Exp()
    ref = lie.so3_exp(phi.astype(np.float64))
    assert np.abs(q.tensor().cpu().numpy() - ref).max() < 1e-6
    back = q.Log().tensor().cpu().numpy()
    ref_back = lie.so3_log(ref)
    assert np.abs(back - ref_back).max() < 5e-6


def test_broadcast_group_times_points_and_groups():
    X = pp.SE3(torch.tensor(_rand_se3(1, 4), device=DEV))                # (1,7)
    Y = pp.SE3(torch.tensor(_rand_se3(9, 5), device=DEV))                # (9,7)
    Z = X @ Y                                                             # pvgo.py:118 target @ source.Inv() @ nodes
    assert Z.shape == (9, 7)
    ref = lie.se3_mul(X.tensor().cpu().numpy().astype(np.float64), Y.tensor().cpu().numpy().astype(np.float64))
    assert np.abs(Z.tensor().cpu().numpy() - ref).max() < 2e-5
    v = torch.randn(9, 3, device=DEV)
    out = X.rotation()[0] @ v                                              # pvgo.py:117 SO3 (4,) @ (N,3)
    assert out.shape == (9, 3)


def _num_grad(fun, X, eps=1e-3):
    """d sum(w * fun(Exp(d) X)) / d d  by central differences in float64 (oracle)."""
    g = np.zeros((X.shape[0], 6))
    for k in range(6):
        d = np.zeros((X.shape[0], 6)); d[:, k] = eps
        g[:, k] = (fun(lie.se3_retract(X, d)) - fun(lie.se3_retract(X, -d))) / (2 * eps)
    return g


def test_backward_matches_left_tangent_finite_differences():
    Xn, Yn = _rand_se3(33, 6, 0.7).astype(np.float64), _rand_se3(33, 7, 0.7).astype(np.float64)
    w6 = np.random.default_rng(8).standard_normal((33, 6))
    # f(X) = w . Log(Y^-1 X^-1 ... ) chain as in pvgo.py:72:  e = Log(P^-1 C)
    x = torch.tensor(Xn, dtype=torch.float32, device=DEV, requires_grad=True)
    y = torch.tensor(Yn, dtype=torch.float32, device=DEV)
    out = (pp.SE3(x).Inv() @ pp.SE3(y)).Log().tensor()
    (out * torch.tensor(w6, dtype=torch.float32, device=DEV)).sum().backward()
    g = x.grad.cpu().numpy()
    assert np.abs(g[:, 6]).max() == 0                                       # 7th slot carries no gradient (A.1)
    ref = _num_grad(lambda X: np.sum(w6 * lie.se3_log(lie.se3_mul(lie.se3_inv(X), Yn)), 1), Xn)
    assert np.abs(g[:, :6] - ref).max() < 2e-3 * max(1.0, np.abs(ref).max())
    # Mul second argument + Act + Exp
    x2 = torch.tensor(Xn, dtype=torch.float32, device=DEV, requires_grad=True)
    p = torch.tensor(np.random.default_rng(9).standard_normal((33, 3)), dtype=torch.float32, device=DEV)
    w3 = np.random.default_rng(10).standard_normal((33, 3))
    out = (pp.SE3(y) @ pp.SE3(x2)) @ p
    (out * torch.tensor(w3, dtype=torch.float32, device=DEV)).sum().backward()
    ref = _num_grad(lambda X: np.sum(w3 * lie.se3_act(lie.se3_mul(Yn, X), p.cpu().numpy().astype(np.float64)), 1), Xn)
    assert np.abs(x2.grad.cpu().numpy()[:, :6] - ref).max() < 2e-3 * max(1.0, np.abs(ref).max())
    xi = torch.tensor(lie.se3_log(Xn), dtype=torch.float32, device=DEV, requires_grad=True)
    out = pp.se3(xi).Exp() @ p
    (out * torch.tensor(w3, dtype=torch.float32, device=DEV)).sum().backward()
    xin = lie.se3_log(Xn)
    gref = np.zeros_like(xin)
    for k in range(6):
        d = np.zeros_like(xin); d[:, k] = 1e-4
        fp = np.sum(w3 * lie.se3_act(lie.se3_exp(xin + d), p.cpu().numpy().astype(np.float64)), 1)
        fm = np.sum(w3 * lie.se3_act(lie.se3_exp(xin - d), p.cpu().numpy().astype(np.float64)), 1)
        gref[:, k] = (fp - fm) / 2e-4
    assert np.abs(xi.grad.cpu().numpy() - gref).max() < 2e-3 * max(1.0, np.abs(gref).max())


def test_add__is_left_retraction():
    X = _rand_se3(5, 11)
    d = np.random.default_rng(12).standard_normal((5, 7)).astype(np.float32) * 0.1
    x = pp.SE3(torch.tensor(X, device=DEV))
    x.add_(torch.tensor(d, device=DEV))
    ref = lie.se3_retract(X.astype(np.float64), d.astype(np.float64))
    assert np.abs(x.tensor().cpu().numpy() - ref).max() < 2e-5


@pytest.mark.parametrize('dev', ['cuda:0', 'cpu'])
def test_trajectory_chaining_helpers(dev):
    """Datasets/transformation.py:100-124 semantics (train.py:220,225,240,264) on the scan / batched kernels."""
    from islam_b200.transformation import motion2pose_pypose, pose2motion_pypose, tartan2kitti_pypose, cvtSE3_pypose
    n = 3000
    M = (lie.se3_exp(np.random.default_rng(20).standard_normal((n, 6)) * np.array([.3, .3, .3, .05, .05, .05]))).astype(np.float32)
    T0 = _rand_se3(1, 21)[0]
    poses = motion2pose_pypose(pp.SE3(torch.tensor(M, device=dev)), pp.SE3(torch.tensor(T0, device=dev)))
    assert poses.shape == (n + 1, 7) and poses.ltype is pp.SE3_type
    ref = [T0.astype(np.float64)]
    for m in M.astype(np.float64):
        ref.append(lie.se3_mul(ref[-1], m))
    ref = np.stack(ref)
    got = poses.tensor().cpu().numpy().astype(np.float64)
    assert np.abs(got[:, :3] - ref[:, :3]).max() < 1e-4 * max(1.0, np.abs(ref[:, :3]).max())
    assert np.abs(lie.quat_canon(got[:, 3:]) - lie.quat_canon(ref[:, 3:])).max() < 1e-5
    back = pose2motion_pypose(poses)
    assert back.shape == (n, 7)
    d = lie.se3_log(lie.se3_mul(lie.se3_inv(M.astype(np.float64)), back.tensor().cpu().numpy().astype(np.float64)))
    assert np.abs(d).max() < 2e-3                                      # float32 poses ~100 m away
    k = tartan2kitti_pypose(torch.tensor(lie.se3_log(M[:5].astype(np.float64)), dtype=torch.float32, device=dev))
    assert k.shape == (5, 7)
    left = pp.cumprod(pp.SE3(torch.tensor(M[:50], device=dev)), left=True).tensor().cpu().numpy()
    r = M[0].astype(np.float64)
    for m in M[1:50].astype(np.float64):
        r = lie.se3_mul(m, r)
    assert np.abs(left[-1] - r).max() < 1e-4
    assert cvtSE3_pypose(pp.SE3(torch.tensor(M[:2], device=dev))).ltype is pp.SE3_type
