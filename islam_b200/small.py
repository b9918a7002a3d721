"""Host side of the small-graph fast path (csrc/small.cuh): `run_pvgo` on the window sizes train.py uses (run_kitti.sh:8:
9 poses) as ONE kernel launch, one packed host->device copy and one packed device->host copy; `run_pvgo_batch` runs many
windows of identical structure at once, one CTA per window.

The general path (solver.PVGOSolver) stays the implementation for everything else: more than 16 poses / 128 edges, the
reprojection factor, target='imu', multi-GPU."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import IslamError, LMParams, LMState

_RUNNERS = {}
_STATE_FLOATS = (C.sizeof(LMState) + 3) // 4


def _plain(t):
    if isinstance(t, torch.Tensor):
        return t.as_subclass(torch.Tensor) if type(t) is not torch.Tensor else t
    return torch.as_tensor(np.asarray(t))


class SmallPVGO:
    """B windows of N poses and E edges (one shared edge list) on one device."""

    def __init__(self, N, links, device, B=1):
        self.L = _lib.lib()
        self.device = torch.device(device)
        self.N, self.M, self.B = int(N), int(N) - 1, int(B)
        self.links_np = np.ascontiguousarray(np.asarray(links), dtype=np.int64).reshape(-1, 2)
        self.E = int(self.links_np.shape[0])
        if not self.L.islam_pvgo_small_supported(self.N, self.E):
            raise IslamError(f'the small-graph path covers N <= 16 poses and E <= 128 edges (got N={self.N}, E={self.E})')
        self.links_dev = torch.as_tensor(self.links_np.astype(np.int32)).to(self.device)
        N_, E_, M_ = self.N, self.E, self.M
        # packed input layout per call: every field window-major (B, ...)
        self.fields = [('nodes0', N_ * 7), ('vels0', N_ * 3), ('Z', E_ * 7), ('drot', M_ * 4), ('dtrans', M_ * 3), ('dvel', M_ * 3),
                       ('dt', M_)]
        self.in_off, off = {}, 0
        for k, n in self.fields:
            self.in_off[k] = (off, n * self.B)
            off += n * self.B
        self.n_in = off
        self.out_fields = [('state', _STATE_FLOATS + (_STATE_FLOATS & 1)), ('nodes', N_ * 7), ('vels', N_ * 3), ('tl', E_), ('rl', E_),
                           ('gt', 6 * E_), ('gr', 6 * E_)]
        self.out_off, off = {}, 0
        for k, n in self.out_fields:
            self.out_off[k] = (off, n * self.B)
            off += n * self.B
        self.n_out_host = self.out_off['gt'][0]                 # gradients stay on the device
        self.n_out = off
        self.h_in = torch.empty(self.n_in, dtype=torch.float32, pin_memory=True)
        self.d_in = torch.empty(self.n_in, dtype=torch.float32, device=self.device)
        self.d_out = torch.empty(self.n_out, dtype=torch.float32, device=self.device)
        self.h_out = torch.empty(self.n_out_host, dtype=torch.float32, pin_memory=True)
        self.params = LMParams()
        self.L.islam_lm_default_params(C.byref(self.params))

    def _seg(self, buf, table, k):
        o, n = table[k]
        return buf[o:o + n]

    def run(self, args, loss_weight, vo_P=None, with_grad=False, **lm):
        """args: dict of the seven input fields (tensors of B windows, host or device).  Returns (state list, nodes, vels, tl, rl,
        gt, gr): nodes / vels as pinned host tensors, losses (and gradients) as device tensors."""
        dev = self.device
        ptrs = {}
        staged = False
        for k, n in self.fields:
            t = _plain(args[k]).detach()
            if t.numel() != n * self.B:
                raise IslamError(f'{k}: expected {n * self.B} numbers, got {t.numel()}')
            if t.is_cuda:
                t = t.to(dtype=torch.float32).contiguous()
                ptrs[k] = (t.data_ptr(), t)
            else:
                self._seg(self.h_in, self.in_off, k).copy_(t.reshape(-1))
                ptrs[k] = (self._seg(self.d_in, self.in_off, k).data_ptr(), None)
                staged = True
        if staged:
            self.d_in.copy_(self.h_in, non_blocking=True)           # ONE host -> device copy
        for k, v in lm.items():
            setattr(self.params, k, v)
        w = (C.c_double * 4)(*[float(x) ** 2 for x in loss_weight[:4]])
        P = None
        if vo_P is not None:
            P = _plain(vo_P).detach().to(device=dev, dtype=torch.float32).contiguous()
        seg = lambda k: self._seg(self.d_out, self.out_off, k)
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(self.L.islam_pvgo_small_run(
                self.B, self.N, self.E, p(self.links_dev), self.links_np.ctypes.data,
                *[C.c_void_p(ptrs[k][0]) for k, _ in self.fields], C.byref(w), C.byref(self.params),
                p(seg('nodes')), p(seg('vels')), p(seg('state')), p(P) if P is not None else C.c_void_p(0),
                p(seg('tl')), p(seg('rl')), p(seg('gt')) if with_grad else C.c_void_p(0), p(seg('gr')) if with_grad else C.c_void_p(0),
                stream), 'islam_pvgo_small_run')
        self.h_out.copy_(self.d_out[:self.n_out_host], non_blocking=True)     # ONE device -> host copy
        torch.cuda.current_stream(dev).synchronize()
        so, sn = self.out_off['state']
        raw = self.h_out[so:so + sn].numpy().tobytes()
        per = (sn // self.B) * 4
        states = [LMState.from_buffer_copy(raw[b * per:b * per + C.sizeof(LMState)]) for b in range(self.B)]
        host = lambda k, shape: self._seg(self.h_out, self.out_off, k).reshape(shape).clone()
        nodes = host('nodes', (self.B, self.N, 7))
        vels = host('vels', (self.B, self.N, 3))
        tl = seg('tl').reshape(self.B, self.E).clone()
        rl = seg('rl').reshape(self.B, self.E).clone()
        gt = seg('gt').reshape(self.B, self.E, 6).clone() if with_grad else None
        gr = seg('gr').reshape(self.B, self.E, 6).clone() if with_grad else None
        return states, nodes, vels, tl, rl, gt, gr


def get_runner(N, links, device, B=1):
    links_np = np.ascontiguousarray(_plain(links).detach().cpu().numpy(), dtype=np.int64).reshape(-1, 2)
    dev = torch.device(device)
    if dev.type == 'cuda' and dev.index is None:
        dev = torch.device('cuda', torch.cuda.current_device())
    key = (int(N), int(B), str(dev), links_np.tobytes())
    r = _RUNNERS.get(key)
    if r is None:
        if len(_RUNNERS) >= 16:
            _RUNNERS.pop(next(iter(_RUNNERS)))
        r = _RUNNERS[key] = SmallPVGO(N, links_np, dev, B)
    return r


class PrecomputedVoLoss(torch.autograd.Function):
    """vo_loss whose value AND gradient came out of the fused launch: backward only scales and pads (A.1: left tangent, 6 -> 7)."""

    @staticmethod
    def forward(ctx, P, tl, rl, gt, gr):
        ctx.save_for_backward(gt, gr)
        ctx.dev = P.device
        return tl.clone(), rl.clone()

    @staticmethod
    def backward(ctx, g_tl, g_rl):
        gt, gr = ctx.saved_tensors
        g = g_tl.to(gt.device).unsqueeze(-1) * gt + g_rl.to(gr.device).unsqueeze(-1) * gr
        g = torch.cat([g, torch.zeros_like(g[..., :1])], dim=-1)
        return g.to(ctx.dev), None, None, None, None
