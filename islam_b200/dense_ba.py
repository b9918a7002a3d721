"""Mirror of the one function of /root/reference/dense_ba.py that sits immediately before the PVGO back-end:
`scale_from_disp_flow` (dense_ba.py:88-176, called per sample at TartanVO.py:159-171).

Same arguments, same four return values.  `scale_from_disp_flow_batch` does the whole batch of TartanVO.py's Python loop
in one fused kernel launch.  There is no CPU path."""
import ctypes as C

import torch

from . import _lib
from ._lib import IslamError
from .pvgo import _plain


def _quat_rot(q, p):
    v, w = q[..., :3], q[..., 3:4]
    t = 2.0 * torch.linalg.cross(v, p)
    return p + w * t + torch.linalg.cross(v, t)


class _ScaleFn(torch.autograd.Function):
    """s = sum(M w) / sum(M M) with autograd into `motion` (dense_ba.py:144-166 runs with grad enabled at TartanVO.py:92, and
    TartanVO.py:181 feeds the scale back into the pose).  The per-pixel sums of the backward pass come out of the SAME fused
    kernel pass as the value (grad_sums, include/islam_pvgo.h); what remains here is the 7-number chain rule per sample, in
    PyPose's convention: T.Inv() is a LieTensor op (left-tangent gradient, -Ad(T^-1)^T), .rotation() keeps the tangent,
    .translation() is a plain slice whose raw gradient lands in the tau slots."""

    @staticmethod
    def forward(ctx, motion, call):
        scale, sums = call(motion.detach(), True)
        ctx.save_for_backward(motion.detach(), sums)
        ctx.intr = call.intr
        return scale

    @staticmethod
    def backward(ctx, g_s):
        mo, sums = ctx.saved_tensors
        mo = mo.double()
        fx, fy, cx, cy = (ctx.intr[:, k].double() for k in range(4))
        num, den = sums[:, 0], sums[:, 1]
        s = num / den
        ds_da = (sums[:, 2:5] - s.unsqueeze(-1) * sums[:, 5:8]) / den.unsqueeze(-1)
        ds_dr = sums[:, 8:11] / den.unsqueeze(-1)
        t, q = mo[:, :3], mo[:, 3:7]
        qi = q * torch.tensor([-1.0, -1.0, -1.0, 1.0], dtype=q.dtype, device=q.device)
        t_inv = -_quat_rot(qi, t)
        nrm = t_inv.norm(dim=1, keepdim=True).clamp_min(1e-12)
        n = t_inv / nrm
        ds_dn = torch.stack([fx * ds_da[:, 0], fy * ds_da[:, 1], cx * ds_da[:, 0] + cy * ds_da[:, 1] + ds_da[:, 2]], dim=1)
        g_tau = (ds_dn - n * (n * ds_dn).sum(1, keepdim=True)) / nrm          # raw gradient w.r.t. T.Inv().translation()
        # Y = T^-1: grad_T = -Ad(T^-1)^T [g_tau; g_phi] = -[R g_tau ; R (g_phi - t_inv x g_tau)]
        g6 = -torch.cat([_quat_rot(q, g_tau), _quat_rot(q, ds_dr - torch.linalg.cross(t_inv, g_tau))], dim=1)
        g7 = torch.cat([g6, torch.zeros_like(g6[:, :1])], dim=1) * g_s.double().reshape(-1, 1)
        return g7.to(torch.float32), None


def _f(t, dev, shape=None):
    t = _plain(t).detach().to(device=dev, dtype=torch.float32).contiguous()
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise IslamError(f'expected shape {tuple(shape)}, got {tuple(t.shape)}')
    return t


def scale_from_disp_flow_batch(disp, flow, motion, intrinsics, baseline, depth=None, mask=None, disp_th=1.0, device=None):
    """disp (B,H,W) [ignored when depth is given], flow (B,2,H,W), motion (B,7) SE3, intrinsics (B,4) = fx,fy,cx,cy,
    baseline (B,), depth (B,H,W) | None, mask (B,H,W) bool | None, disp_th float | (B,).
    Returns scale (B,), z (B,H,W), mask (B,H,W) bool, depth_mask (B,H,W) bool, mask_count (B,) int32 — on the device.
    `scale` carries autograd history into `motion` when that requires grad (as the reference's does)."""
    flow_t = _plain(flow)
    dev = torch.device(device) if device is not None else (flow_t.device if flow_t.is_cuda else torch.device('cuda', torch.cuda.current_device()))
    if dev.type != 'cuda':
        raise IslamError('scale_from_disp_flow needs a CUDA device: there is no CPU fallback')
    flow_t = _f(flow_t, dev)
    B, two, H, W = flow_t.shape
    if two != 2:
        raise IslamError('flow must be (B, 2, H, W)')
    mo_in = _plain(motion)
    if mo_in.shape[-1] == 6:                    # dense_ba.py:92-95: an se3 input goes through Exp first (differentiable)
        from .pypose_compat import _ops
        mo_in = _ops.ExpFn.apply(mo_in.to(device=dev, dtype=torch.float32), _ops.SE3)
    elif mo_in.shape[-1] != 7:
        raise IslamError('motion must be SE3 (7 numbers) or se3 (6 numbers)')
    mo_in = mo_in.to(device=dev, dtype=torch.float32).reshape(B, 7)
    intr = _f(intrinsics, dev, (B, 4))
    bl = _f(torch.as_tensor(baseline).reshape(-1), dev, (B,))
    th = torch.as_tensor(disp_th, dtype=torch.float32).reshape(-1)
    th = _f(th.expand(B) if th.numel() == 1 else th, dev, (B,))
    d = _f(disp, dev, (B, H, W)) if depth is None else None
    dep = _f(depth, dev, (B, H, W)) if depth is not None else None
    m_in = _plain(mask).to(device=dev, dtype=torch.uint8).contiguous() if mask is not None else None
    if m_in is not None and tuple(m_in.shape) != (B, H, W):
        raise IslamError(f'expected mask shape {(B, H, W)}, got {tuple(m_in.shape)}')
    L = _lib.lib()
    ws = torch.empty(int(L.islam_scale_workspace_bytes(B, H, W)), dtype=torch.uint8, device=dev)
    z = torch.empty(B, H, W, dtype=torch.float32, device=dev)
    m_out = torch.empty(B, H, W, dtype=torch.uint8, device=dev)
    dm_out = torch.empty(B, H, W, dtype=torch.uint8, device=dev)
    cnt = torch.empty(B, dtype=torch.int32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    def call(mo, want_grad):
        mo = mo.contiguous()
        scale = torch.empty(B, dtype=torch.float32, device=dev)
        sums = torch.empty(B, 11, dtype=torch.float64, device=dev) if want_grad else None
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(L.islam_scale_from_disp_flow(p(d), p(flow_t), p(mo), p(intr), p(bl), p(dep), p(m_in), p(th), B, H, W, p(scale),
                                                    p(z), p(m_out), p(dm_out), p(cnt), p(sums), p(ws), st), 'islam_scale_from_disp_flow')
        return (scale, sums) if want_grad else scale
    call.intr = intr

    if mo_in.requires_grad and torch.is_grad_enabled():
        scale = _ScaleFn.apply(mo_in, call)
    else:
        scale = call(mo_in.detach(), False)
    return scale, z, m_out.bool(), dm_out.bool(), cnt


def scale_from_disp_flow(disp, flow, motion, fx, fy, cx, cy, baseline, depth=None, mask=None, disp_th=1):
    """dense_ba.py:88 — one sample: disp (H,W), flow (2,H,W), motion SE3 (7,).  Returns (s (1,), z (H,W), mask, depth_mask)."""
    mo = _plain(motion)
    dev = mo.device if mo.is_cuda else None
    intr = torch.tensor([[float(fx), float(fy), float(cx), float(cy)]], dtype=torch.float32)
    un = lambda t: None if t is None else _plain(t).unsqueeze(0)
    s, z, m, dm, cnt = scale_from_disp_flow_batch(un(disp), un(flow), mo.reshape(1, -1), intr, torch.tensor([float(baseline)]),
                                                 depth=un(depth), mask=un(mask), disp_th=float(disp_th), device=dev)
    if int(cnt.item()) < 500:
        print('Warning! mask contains too less points!', int(cnt.item()))          # dense_ba.py:134-135
    return s.view(1), z[0], m[0], dm[0]


def pixel2point(pixels, depth, intrinsics):
    """dense_ba.py:9-64: pixels (...,N,2) with depth (...,N) -> camera-frame points (...,N,3)."""
    assert pixels.size(-1) == 2 and depth.size(-1) == pixels.size(-2)
    fx, fy = intrinsics[..., 0, 0], intrinsics[..., 1, 1]
    cx, cy = intrinsics[..., 0, 2], intrinsics[..., 1, 2]
    f = torch.stack([fx, fy], dim=-1).unsqueeze(-2)
    c = torch.stack([cx, cy], dim=-1).unsqueeze(-2)
    xy = (pixels - c) / f * depth.unsqueeze(-1)
    return torch.cat([xy, depth.unsqueeze(-1)], dim=-1)


class SparseReprojectionLoss:
    """Mirror of dense_ba.py:276-305, the object `run_pvgo(..., reproj=...)` takes (pvgo.py:53-61,130-165): N sparse points per
    consecutive pair, their 3-D positions in the first camera and their pixel targets (flow + pixel) in the second.
    As a PVGO factor it is consumed by the fused kernels (csrc/linearize.cuh rp_block); called directly it evaluates the
    residual (batch, N, 2) on the shim's LieTensor ops like the reference's __call__."""

    def __init__(self, points2d, depth, flow, fx, fy, cx, cy, rgb2imu_pose, device='cuda:0'):
        assert len(flow.shape) == 4 and len(depth.shape) == 3 and len(points2d.shape) == 3
        bs, N = points2d.shape[:2]
        idx = torch.cat([torch.arange(0, bs).repeat_interleave(N).view(bs, N, 1), _plain(points2d)[..., (1, 0)].cpu()], dim=-1).to(int)
        points2d = _plain(points2d).to(device)
        idx = idx.to(device)
        depth, flow = _plain(depth).to(device), _plain(flow).to(device)
        self.K = torch.tensor([fx, 0, cx, 0, fy, cy, 0, 0, 1], dtype=torch.float32).view(3, 3).to(device)
        self.point3d = pixel2point(points2d, depth[idx[..., 0], idx[..., 1], idx[..., 2]].view(bs, N), self.K).to(device)
        self.target = (flow.permute(0, 2, 3, 1)[idx[..., 0], idx[..., 1], idx[..., 2], :].view(bs, N, 2) + points2d).to(device)
        self.N = N
        self.rgb2imu_pose = rgb2imu_pose.to(device)

    def __call__(self, motion):
        from .pypose_compat.function.geometry import reprojerr
        T = self.rgb2imu_pose.Inv() @ motion @ self.rgb2imu_pose
        return reprojerr(self.point3d, self.target, self.K, T.Inv(), reduction='none')
