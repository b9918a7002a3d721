"""Golden vectors for SURVEY 8f rank 4, `scale_from_disp_flow` (/root/reference/dense_ba.py:88-176), produced by running
THE REFERENCE FUNCTION ITSELF here in the build container (run once; /root/reference does not exist on the GPU box):

    python tests/golden/make_scale_golden.py        ->  tests/golden/scale_golden.npz

PyPose is absent, so the five LieTensor calls the function makes (pp.SE3(x), .Inv(), .rotation(), .translation(),
.tensor(), SO3 @ points) are served by the ~30-line stand-in below (quaternion xyzw arithmetic); everything
else — masks, back-projection, the linear system, the least-squares scale — is the reference's own torch code, float32 on
the CPU.  Scenes: a synthetic static world seen by a translating + rotating camera (exact flow, so the recovered scale must
equal the true baseline ratio), the same with noisy flow / an edge mask / the depth-input branch, and pure-noise inputs.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'


# ---- minimal PyPose stand-in (only what dense_ba.scale_from_disp_flow touches) --------------------------------------------
def _qrot(q, p):
    v, w = q[..., :3], q[..., 3:]
    t = 2 * torch.cross(v.expand_as(p), p, dim=-1)
    return p + w * t + torch.cross(v.expand_as(p), t, dim=-1)


class _SO3:
    def __init__(self, q): self.q = q
    def unsqueeze(self, d): return _SO3(self.q.unsqueeze(d))
    def tensor(self): return self.q
    def __matmul__(self, p): return _qrot(self.q, p)


class _SE3:
    def __init__(self, x): self.x = torch.as_tensor(x)
    def Inv(self):
        t, q = self.x[..., :3], self.x[..., 3:]
        qi = torch.cat([-q[..., :3], q[..., 3:]], -1)
        return _SE3(torch.cat([-_qrot(qi, t), qi], -1))
    def rotation(self): return _SO3(self.x[..., 3:])
    def translation(self): return self.x[..., :3]
    def tensor(self): return self.x


def _install_stub():
    pp = types.ModuleType('pypose')
    pp.SE3 = _SE3
    geo = types.ModuleType('pypose.function.geometry')
    geo.reprojerr = geo.point2pixel = None
    fn = types.ModuleType('pypose.function')
    fn.geometry = geo
    pp.function = fn
    sys.modules.update({'pypose': pp, 'pypose.function': fn, 'pypose.function.geometry': geo})


def scene(rng, H=64, W=96, kind="exact"):
    """A static cloud at depths 4..30 m, camera moves by T (camera frame k -> k+1, as TartanVO reports it)."""
    fx = fy = 80.0 + 10 * rng.random(); cx, cy = W / 2 - 0.5, H / 2 - 0.5
    baseline = 0.5
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    z = 4 + 26 * rng.random((H, W))
    z = z * (0.7 + 0.3 * np.sin(u / 17.0) * np.cos(v / 11.0))            # smooth-ish structure
    P = np.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], -1)
    # motion = pose of frame k+1 in frame k; scale_from_disp_flow receives it with a UNIT-norm translation direction
    ang = rng.standard_normal(3) * 0.02
    th = np.linalg.norm(ang)
    q = np.concatenate([np.sin(th / 2) * ang / th, [np.cos(th / 2)]])
    t_true = np.array([0.05, -0.02, 0.35]) * (0.5 + rng.random())
    T = torch.tensor(np.concatenate([t_true, q]))
    Ti = _SE3(T).Inv()
    P1 = _qrot(Ti.x[3:], torch.tensor(P)) + Ti.x[:3]                     # points in frame k+1
    uv1 = torch.stack([fx * P1[..., 0] / P1[..., 2] + cx, fy * P1[..., 1] / P1[..., 2] + cy], 0).numpy()
    flow = uv1 - np.stack([u, v], 0)
    disp = fx * baseline / z
    depth = None
    mask = None
    motion = np.concatenate([t_true / np.linalg.norm(t_true), q])          # direction only: the scale is what is estimated
    if kind == 'noisy':
        flow = flow + 0.3 * rng.standard_normal(flow.shape)
        disp = disp * (1 + 0.02 * rng.standard_normal(disp.shape))
    if kind == 'masked':
        mask = rng.random((H, W)) > 0.6
        flow[:, :5, :] = 0.0                                              # zero flow rows are rejected (norm > 0)
    if kind == 'depth':
        depth = z * (1 + 0.01 * rng.standard_normal(z.shape))
        depth[::7, ::5] = -1.0                                            # invalid depths
    if kind == 'random':
        flow = 20 * rng.standard_normal(flow.shape)
        disp = np.abs(6 * rng.standard_normal(disp.shape))
    return dict(disp=disp.astype(np.float32), flow=flow.astype(np.float32), motion=motion.astype(np.float32),
                intr=np.array([fx, fy, cx, cy], np.float32), baseline=np.float32(baseline),
                depth=None if depth is None else depth.astype(np.float32), mask=mask,
                disp_th=np.float32(5.0 if kind == 'random' else 1.0), s_true=np.float32(np.linalg.norm(t_true)))


def main():
    _install_stub()
    sys.path.insert(0, REF)
    import dense_ba                                                        # the reference module, unmodified
    rng = np.random.default_rng(11)
    out = {}
    kinds = ['exact', 'noisy', 'masked', 'depth', 'random', 'exact']
    for k, kind in enumerate(kinds):
        sc = scene(rng, kind=kind)
        t = lambda a: None if a is None else torch.as_tensor(a)
        s, z, m, dm = dense_ba.scale_from_disp_flow(t(sc['disp']), t(sc['flow']), t(sc['motion']), *[float(x) for x in sc['intr']],
                                                    float(sc['baseline']), depth=t(sc['depth']), mask=t(sc['mask']),
                                                    disp_th=float(sc['disp_th']))
        for name in ('disp', 'flow', 'motion', 'intr', 'baseline', 'disp_th', 's_true'):
            out[f'{k}_{name}'] = sc[name]
        out[f'{k}_kind'] = np.array(kind)
        out[f'{k}_has_depth'] = np.array(sc['depth'] is not None)
        out[f'{k}_has_mask'] = np.array(sc['mask'] is not None)
        if sc['depth'] is not None: out[f'{k}_depth'] = sc['depth']
        if sc['mask'] is not None: out[f'{k}_mask'] = sc['mask']
        out[f'{k}_ref_s'] = s.numpy(); out[f'{k}_ref_z'] = z.numpy(); out[f'{k}_ref_mask'] = m.numpy(); out[f'{k}_ref_dmask'] = dm.numpy()
        print(kind, 'reference scale', float(s), 'true', float(sc['s_true']), 'mask', int(m.sum()))
    out['n'] = np.array(len(kinds))
    np.savez_compressed(os.path.join(HERE, 'scale_golden.npz'), **out)


if __name__ == '__main__':
    main()
