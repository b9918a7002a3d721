"""The reprojection-factor oracle (oracle/reproj_oracle.py) pinned against outputs of the reference class
`SparseReprojectionLoss` (tests/golden/reproj_golden.npz, made by make_reproj_golden.py), and its Jacobian against finite
differences of the left perturbation X <- Exp(d) X."""
import os

import numpy as np
import pytest

from oracle import lie, reproj_oracle as ro

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reproj_golden.npz'))


def _rp(name):
    return dict(point3d=G[f'{name}_point3d'], target=G[f'{name}_target'], K=G[f'{name}_K'], rgb2imu=G[f'{name}_rgb2imu'])


@pytest.mark.parametrize('name', [str(c) for c in G['cases']])
def test_residual_matches_reference_class(name):
    want = G[f'{name}_err']
    got32 = ro.residual(G[f'{name}_nodes'].astype(np.float32), _rp(name))
    got64 = ro.residual(G[f'{name}_nodes'].astype(np.float64), _rp(name))
    assert got32.shape == want.shape
    scale = np.abs(want).max()
    assert np.abs(got64 - want).max() <= 2e-5 * scale          # float32 reference vs float64 restatement
    assert np.abs(got32 - want).max() <= 2e-5 * scale
    # row 0 is the overwritten motion (pvgo.py:57): a constant, however the nodes move
    moved = G[f'{name}_nodes'].astype(np.float64).copy()
    moved[:, :3] += 0.3
    assert np.array_equal(ro.residual(moved, _rp(name))[0], got64[0])


def test_jacobian_matches_finite_differences():
    name = str(G['cases'][0])
    nodes = G[f'{name}_nodes'].astype(np.float64)
    rp = _rp(name)
    J = ro.jacobian(nodes, rp)
    M = nodes.shape[0] - 1
    eps = 1e-6
    for i in (0, 1, 3, M - 1):
        for k in range(6):
            d = np.zeros(6); d[k] = eps
            for which, sign in ((i, 1.0), (i + 1, -1.0)):              # d r_i / d delta_i = J_i, d r_i / d delta_{i+1} = -J_i
                p, m_ = nodes.copy(), nodes.copy()
                p[which] = lie.se3_retract(nodes[which][None], d[None])[0]
                m_[which] = lie.se3_retract(nodes[which][None], -d[None])[0]
                fd = (ro.residual(p, rp)[i] - ro.residual(m_, rp)[i]) / (2 * eps)
                want = sign * J[i][:, k]
                if i == 0:
                    assert np.abs(fd).max() == 0 and np.abs(want).max() == 0      # constant row
                else:
                    assert np.abs(fd - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), (i, k, which)
