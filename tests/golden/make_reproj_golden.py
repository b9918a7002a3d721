"""Golden vectors for the optional reprojection factor (/root/reference/pvgo.py:53-61,
/root/reference/dense_ba.py:276-305), produced by running THE REFERENCE CLASS `SparseReprojectionLoss` ITSELF — constructor
and __call__ — here in the build container (/root/reference does not exist on the GPU box):

    python tests/golden/make_reproj_golden.py       ->  tests/golden/reproj_golden.npz

PyPose is absent, so the LieTensor calls on this path (pp.SE3(x), .Inv(), @ between poses, pose @ points with broadcasting,
.to, .unsqueeze, slicing, item assignment) and `pypose.function.geometry.reprojerr / point2pixel` are served by the stand-in
below: quaternion xyzw arithmetic exactly as SURVEY.md A.1 states it (no normalisation anywhere — the overwritten motion[0]
of pvgo.py:57 is not a unit quaternion), and point2pixel = homo2cart(points @ K^T) after the extrinsic action.  Everything
else (pixel2point, the gathers that build point3d / target, the composition rgb2imu^-1 motion rgb2imu, the residual) is the
reference's own code, float32 on the CPU.  The motion fed to the loss is built as pvgo.py:54-57 does.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)


def _qrot(q, p):
    v, w = q[..., :3], q[..., 3:]
    v, p = torch.broadcast_tensors(v, p)
    t = 2 * torch.cross(v, p, dim=-1)
    return p + w * t + torch.cross(v, t, dim=-1)


def _qmul(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], -1)


class SE3:
    """pp.SE3 stand-in: a (…,7) tensor [t, q xyzw]."""
    def __init__(self, x): self.x = torch.as_tensor(x)
    def Inv(self):
        t, q = self.x[..., :3], self.x[..., 3:]
        qi = torch.cat([-q[..., :3], q[..., 3:]], -1)
        return SE3(torch.cat([-_qrot(qi, t), qi], -1))
    def __matmul__(self, o):
        if isinstance(o, SE3):
            a, b = torch.broadcast_tensors(self.x, o.x)
            return SE3(torch.cat([a[..., :3] + _qrot(a[..., 3:], b[..., :3]), _qmul(a[..., 3:], b[..., 3:])], -1))
        return _qrot(self.x[..., 3:], o) + self.x[..., :3]
    def to(self, *a, **k): return SE3(self.x.to(*a, **k))
    def unsqueeze(self, d): return SE3(self.x.unsqueeze(d))
    def __getitem__(self, i): return SE3(self.x[i])
    def __setitem__(self, i, v): self.x[i] = v
    def tensor(self): return self.x


def point2pixel(points, intrinsics, extrinsics=None):
    if extrinsics is not None:
        points = extrinsics.unsqueeze(-2) @ points
    h = points @ intrinsics.mT
    tiny = torch.finfo(h.dtype).tiny
    den = h[..., -1:].abs().clamp_(min=tiny)
    den = torch.where(h[..., -1:] >= 0, den, -den)
    return h[..., :-1] / den


def reprojerr(points, pixels, intrinsics, extrinsics=None, reduction='none'):
    err = point2pixel(points, intrinsics, extrinsics) - pixels
    return err.norm(dim=-1) if reduction == 'norm' else err


def _install_stub():
    pp = types.ModuleType('pypose')
    pp.SE3 = SE3
    geo = types.ModuleType('pypose.function.geometry')
    geo.reprojerr, geo.point2pixel = reprojerr, point2pixel
    fn = types.ModuleType('pypose.function')
    fn.geometry = geo
    pp.function = fn
    sys.modules.update({'pypose': pp, 'pypose.function': fn, 'pypose.function.geometry': geo})


def main():
    _install_stub()
    sys.path.insert(0, REF)
    import dense_ba                                   # the reference module, unmodified
    from islam_b200 import synth
    out = {}
    cases = [('win9', synth.window(), 16), ('chain20', synth.config3(N=20), 12)]
    for name, g, npts in cases:
        rng = np.random.default_rng(len(name))
        M, H, W = g.N - 1, 16, 24
        fx, fy, cx, cy = 20.0, 21.0, 11.5, 7.5
        depth = torch.tensor(4 + 20 * rng.random((M, H, W)), dtype=torch.float32)
        flow = torch.tensor(2.0 * rng.standard_normal((M, 2, H, W)), dtype=torch.float32)
        pts2d = torch.tensor(np.stack([rng.integers(0, W, (M, npts)), rng.integers(0, H, (M, npts))], -1), dtype=torch.float32)
        C = synth.reproj_data(g, 4)['rgb2imu']
        loss = dense_ba.SparseReprojectionLoss(pts2d, depth, flow, fx, fy, cx, cy, SE3(torch.tensor(C)), device='cpu')
        nodes = torch.tensor(g.init_nodes)
        motion = SE3(nodes[:-1]).Inv() @ SE3(nodes[1:])             # pvgo.py:54-56
        motion[0] = 0.1                                             # pvgo.py:57
        err = loss(motion)                                          # dense_ba.py:299-305
        assert err.shape == (M, npts, 2) and loss.N == npts
        out[f'{name}_nodes'] = g.init_nodes
        out[f'{name}_point3d'] = loss.point3d.numpy()
        out[f'{name}_target'] = loss.target.numpy()
        out[f'{name}_K'] = np.array([fx, fy, cx, cy], np.float32)
        out[f'{name}_rgb2imu'] = C
        out[f'{name}_err'] = err.reshape(M, npts * 2).numpy()       # pvgo.py:59-60
        # the constructor's own products, for the mirror class (islam_b200.dense_ba.SparseReprojectionLoss)
        out[f'{name}_pts2d'] = pts2d.numpy(); out[f'{name}_depth'] = depth.numpy(); out[f'{name}_flow'] = flow.numpy()
    out['cases'] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(HERE, 'reproj_golden.npz'), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
