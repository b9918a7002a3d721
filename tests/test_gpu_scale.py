"""SURVEY 8f rank 4 on the GPU: islam_scale_from_disp_flow (through the C ABI) against the golden outputs of the reference
function and against the float64 oracle, single-sample and batched."""
import os

import numpy as np
import pytest
import torch

from islam_b200 import dense_ba
from oracle import dense_ba_oracle as dbo

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'scale_golden.npz'))


def _case(k):
    kw = dict(depth=G[f'{k}_depth'] if bool(G[f'{k}_has_depth']) else None, mask=G[f'{k}_mask'] if bool(G[f'{k}_has_mask']) else None,
              disp_th=float(G[f'{k}_disp_th']))
    fx, fy, cx, cy = [float(x) for x in G[f'{k}_intr']]
    return (G[f'{k}_disp'], G[f'{k}_flow'], G[f'{k}_motion'], fx, fy, cx, cy, float(G[f'{k}_baseline'])), kw


@pytest.mark.parametrize('k', range(int(G['n'])))
def test_single_sample_matches_reference_golden(k):
    args, kw = _case(k)
    t = lambda a: None if a is None else torch.as_tensor(a).cuda()
    s, z, m, dm = dense_ba.scale_from_disp_flow(t(args[0]), t(args[1]), t(args[2]), *args[3:], depth=t(kw['depth']), mask=t(kw['mask']),
                                                disp_th=kw['disp_th'])
    assert s.shape == (1,) and z.shape == args[0].shape and m.dtype == torch.bool
    assert np.array_equal(dm.cpu().numpy(), G[f'{k}_ref_dmask'])
    assert np.array_equal(m.cpu().numpy(), G[f'{k}_ref_mask'])
    assert np.abs(z.cpu().numpy() - G[f'{k}_ref_z']).max() <= 1e-5 * np.abs(G[f'{k}_ref_z']).max()
    ref = float(G[f'{k}_ref_s'][0])
    assert abs(float(s) - ref) <= 2e-4 * abs(ref), (float(s), ref)       # float32 reference sums vs float64 accumulation
    so = dbo.scale_from_disp_flow(*args, **kw)[0]
    assert abs(float(s) - so) <= 2e-4 * abs(so)


def test_batch_equals_per_sample_and_is_deterministic():
    ks = [k for k in range(int(G['n'])) if not bool(G[f'{k}_has_depth']) and not bool(G[f'{k}_has_mask'])]
    disp = torch.as_tensor(np.stack([G[f'{k}_disp'] for k in ks])).cuda()
    flow = torch.as_tensor(np.stack([G[f'{k}_flow'] for k in ks])).cuda()
    mo = torch.as_tensor(np.stack([G[f'{k}_motion'] for k in ks])).cuda()
    intr = torch.as_tensor(np.stack([G[f'{k}_intr'] for k in ks])).cuda()
    bl = torch.as_tensor(np.array([G[f'{k}_baseline'] for k in ks])).cuda()
    th = torch.as_tensor(np.array([G[f'{k}_disp_th'] for k in ks])).cuda()
    out1 = dense_ba.scale_from_disp_flow_batch(disp, flow, mo, intr, bl, disp_th=th)
    out2 = dense_ba.scale_from_disp_flow_batch(disp, flow, mo, intr, bl, disp_th=th)
    assert all(torch.equal(a, b) for a, b in zip(out1, out2))                # fixed-order reductions
    for i, k in enumerate(ks):
        assert abs(float(out1[0][i]) - float(G[f'{k}_ref_s'][0])) <= 2e-4 * abs(float(G[f'{k}_ref_s'][0]))
        assert int(out1[4][i]) == int(G[f'{k}_ref_mask'].sum())


def test_full_resolution_property():
    """640 x 448 / 4 = 160 x 112 grid (TartanVO.py:121-122 quarter resolution), batch 8: exact flow -> exact scale."""
    torch.manual_seed(0)
    B, H, W = 8, 112, 160
    fx = fy = 80.0; cx, cy = W / 2 - 0.5, H / 2 - 0.5
    u, v = torch.meshgrid(torch.arange(W, dtype=torch.float64), torch.arange(H, dtype=torch.float64), indexing='xy')
    z = 5 + 20 * torch.rand(B, H, W, dtype=torch.float64)
    t = torch.tensor([0.03, -0.01, 0.4], dtype=torch.float64) * (0.5 + torch.rand(B, 1, dtype=torch.float64))
    P = torch.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], -1)
    P1 = P - t.view(B, 1, 1, 3)                                            # identity rotation: T^-1 p = p - t
    flow = torch.stack([fx * P1[..., 0] / P1[..., 2] + cx - u, fy * P1[..., 1] / P1[..., 2] + cy - v], 1)
    motion = torch.cat([t / t.norm(dim=1, keepdim=True), torch.tensor([[0., 0, 0, 1]]).expand(B, 4).double()], 1)
    s, zz, m, dm, cnt = dense_ba.scale_from_disp_flow_batch((fx * 0.5 / z).float().cuda(), flow.float().cuda(), motion.float().cuda(),
                                                           torch.tensor([[fx, fy, cx, cy]]).expand(B, 4).cuda(), torch.full((B,), 0.5).cuda())
    assert (cnt > 500).all()
    assert torch.allclose(s.cpu().double(), t.norm(dim=1), rtol=2e-4)
