"""Thin torch-facing wrapper over the C ABI (include/islam_pvgo.h): device memory and streams come from PyTorch,
everything numeric happens inside libislam_pvgo.so.  Mirrors the life cycle of /root/reference/pvgo.py:168-197
(PoseVelGraph + pp.optim.LM + StopOnPlateau) for one fixed graph structure."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import IslamError, LMParams, LMState, PvgoDims, PvgoOpts


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _f32(t, device, shape=None):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(t))
    # non_blocking: a pinned host source is copied asynchronously on the current stream (the solver's stream waits for it)
    t = t.detach().to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise IslamError(f'expected shape {tuple(shape)}, got {tuple(t.shape)}')
    return t


class PVGOSolver:
    """One pose-velocity graph on one GPU (or one window of it when n_parts > 1)."""

    def __init__(self, N, links, device='cuda:0', band_max=0, leaf_max=0, pivot_max=0, n_parts=1, part=0):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise IslamError('PVGOSolver needs a CUDA device: the B200 kernels are the only implementation')
        self.L = _lib.lib()
        links = np.ascontiguousarray(np.asarray(links.cpu() if isinstance(links, torch.Tensor) else links),
                                     dtype=np.int64).reshape(-1, 2)
        self.N, self.E, self.M = int(N), int(links.shape[0]), int(N) - 1
        self.links = links
        opts = PvgoOpts(band_max=band_max, leaf_max=leaf_max, pivot_max=pivot_max, n_parts=n_parts, part=part)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.L.islam_pvgo_create(C.byref(self._h), self.N, self.E, links.ctypes.data, C.byref(opts)),
                       'islam_pvgo_create')
        d = PvgoDims()
        _lib.check(self.L.islam_pvgo_get_dims(self._h, C.byref(d)), 'islam_pvgo_get_dims')
        self.dims = d
        self.params = LMParams()
        self.n_reproj = 0
        self.L.islam_lm_default_params(C.byref(self.params))
        # a private non-default stream: the LM loop is CUDA-graph captured, which the legacy stream cannot be
        self.stream = torch.cuda.Stream(device=self.device)

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h:
            try:
                self.L.islam_pvgo_destroy(h)
            except Exception:
                pass

    # ------------------------------------------------------------------------------------------------ helpers
    def _s(self):
        return C.c_void_p(self.stream.cuda_stream)

    def _enter(self):
        cur = torch.cuda.current_stream(self.device)
        if cur != self.stream:                        # (run_pvgo works on the solver's stream: nothing to order)
            self.stream.wait_stream(cur)

    def _exit(self):
        cur = torch.cuda.current_stream(self.device)
        if cur != self.stream:
            cur.wait_stream(self.stream)

    def _new(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    # ------------------------------------------------------------------------------------------------ problem
    def set_problem(self, vo_motions, imu_drots, imu_dtrans, imu_dvels, dts, loss_weight=(1, 1, 1, 1), reproj=None):
        """pvgo.py:125-165: information scalars are loss_weight**2 (VO trans and rot both use loss_weight[0])."""
        dev = self.device
        self._set_reproj(reproj, loss_weight)
        Z = _f32(vo_motions, dev, (self.E, 7))
        dr = _f32(imu_drots, dev, (self.M, 4))
        dp = _f32(imu_dtrans, dev, (self.M, 3))
        dv = _f32(imu_dvels, dev, (self.M, 3))
        dt = _f32(dts, dev).reshape(-1)
        if dt.numel() != self.M:                      # pvgo.py:51 broadcasts vels[:-1] * dts: needs len(dts) == N-1
            raise IslamError(f'dts must have N-1 = {self.M} entries, got {dt.numel()}')
        w = (C.c_double * 4)(float(loss_weight[0]) ** 2, float(loss_weight[1]) ** 2, float(loss_weight[2]) ** 2,
                             float(loss_weight[3]) ** 2)
        self._enter()
        _lib.check(self.L.islam_pvgo_set_problem(self._h, _ptr(Z), _ptr(dr), _ptr(dp), _ptr(dv), _ptr(dt), C.byref(w),
                                                 self._s()), 'islam_pvgo_set_problem')
        for t in (Z, dr, dp, dv, dt):                 # the D2D copies are in flight on our stream: keep the sources alive
            t.record_stream(self.stream)

    def _set_reproj(self, reproj, loss_weight):
        """The optional 5th residual group (pvgo.py:53-61): `reproj` is a SparseReprojectionLoss-shaped object (dense_ba.py:276-305:
        attributes N, point3d (M,N,3), target (M,N,2), K (3,3), rgb2imu_pose SE3), information (loss_weight[4] / N)^2 (pvgo.py:131)."""
        if reproj is None:
            if self.n_reproj:
                self._enter()
                _lib.check(self.L.islam_pvgo_set_reproj(self._h, None, None, 0, None, None, 0.0, self._s()), 'islam_pvgo_set_reproj')
                self.n_reproj = 0
            return
        for name in ('N', 'point3d', 'target', 'K', 'rgb2imu_pose'):
            if not hasattr(reproj, name):
                raise IslamError(f'reproj must look like dense_ba.SparseReprojectionLoss (missing `{name}`); the dense variant '
                                 'cannot be a PVGO factor in the reference either (pvgo.py:131 needs reproj.N)')
        if len(loss_weight) < 5:
            raise IndexError('loss_weight needs a 5th entry when reproj is given (pvgo.py:131)')
        n = int(reproj.N)
        pts = _f32(reproj.point3d, self.device, (self.M, n, 3))
        tgt = _f32(reproj.target, self.device, (self.M, n, 2))
        K = torch.as_tensor(reproj.K).detach().cpu().to(torch.float32).reshape(3, 3)
        intr = (C.c_float * 4)(float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]))
        cal = reproj.rgb2imu_pose
        cal = cal.detach().as_subclass(torch.Tensor) if isinstance(cal, torch.Tensor) else torch.as_tensor(np.asarray(cal))
        c7 = (C.c_float * 7)(*[float(x) for x in cal.cpu().to(torch.float32).reshape(7)])
        self._enter()
        _lib.check(self.L.islam_pvgo_set_reproj(self._h, _ptr(pts), _ptr(tgt), n, C.byref(intr), C.byref(c7),
                                                (float(loss_weight[4]) / n) ** 2, self._s()), 'islam_pvgo_set_reproj')
        for t in (pts, tgt):
            t.record_stream(self.stream)
        self.n_reproj = n

    def set_state(self, nodes, vels):
        n = _f32(nodes, self.device, (self.N, 7))
        v = _f32(vels, self.device, (self.N, 3))
        self._enter()
        _lib.check(self.L.islam_pvgo_set_state(self._h, _ptr(n), _ptr(v), self._s()), 'islam_pvgo_set_state')
        n.record_stream(self.stream)
        v.record_stream(self.stream)

    def get_state(self):
        n, v = self._new(self.N, 7), self._new(self.N, 3)
        _lib.check(self.L.islam_pvgo_get_state(self._h, _ptr(n), _ptr(v), self._s()), 'islam_pvgo_get_state')
        self._exit()
        return n, v

    # ------------------------------------------------------------------------------------------------ family 1
    def linearize(self):
        self._enter()
        _lib.check(self.L.islam_pvgo_linearize(self._h, self._s()), 'islam_pvgo_linearize')

    def residuals(self):
        """(pgerr (E,6), adjvelerr (M,3), imuroterr (M,3), transvelerr (M,3)[, reprojerr (M, 2 N_points)]) — pvgo.py:61-64 order."""
        r = (self._new(self.E, 6), self._new(self.M, 3), self._new(self.M, 3), self._new(self.M, 3))
        _lib.check(self.L.islam_pvgo_get_residuals(self._h, *[_ptr(t) for t in r], self._s()),
                   'islam_pvgo_get_residuals')
        if self.n_reproj:                       # pvgo.py:58-61: a 5th group, (M, 2 N_points)
            rp = self._new(self.M, 2 * self.n_reproj)
            _lib.check(self.L.islam_pvgo_get_reproj_residuals(self._h, _ptr(rp), self._s()), 'islam_pvgo_get_reproj_residuals')
            r = r + (rp,)
        self._exit()
        return r

    def normal_equations(self):
        P = self.dims.P
        Hd = self._new(self.N, 9, 9, dtype=torch.float64)
        Ho = self._new(P, 9, 9, dtype=torch.float64)
        g = self._new(self.N, 9, dtype=torch.float64)
        pairs = self._new(P, 2, dtype=torch.int32)
        _lib.check(self.L.islam_pvgo_get_normal_eq(self._h, _ptr(Hd), _ptr(Ho), _ptr(g), _ptr(pairs), self._s()),
                   'islam_pvgo_get_normal_eq')
        self._exit()
        return Hd, Ho, g, pairs

    # ------------------------------------------------------------------------------------------------ family 2
    def solve(self, diag_scale, lm_min=1e-4, lm_max=1e32):
        D = self._new(self.N, 9, dtype=torch.float64)
        info = C.c_int32(0)
        _lib.check(self.L.islam_pvgo_solve(self._h, float(diag_scale), float(lm_min), float(lm_max), _ptr(D),
                                           C.byref(info), self._s()), 'islam_pvgo_solve')
        self._exit()
        return D, int(info.value)

    # ------------------------------------------------------------------------------------------------ LM
    def lm_reset(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.params, k):
                raise IslamError(f'unknown LM parameter {k}')
            setattr(self.params, k, v)
        self._enter()
        _lib.check(self.L.islam_pvgo_lm_reset(self._h, C.byref(self.params), self._s()), 'islam_pvgo_lm_reset')

    def lm_step(self):
        st = LMState()
        self._enter()
        _lib.check(self.L.islam_pvgo_lm_step(self._h, C.byref(st), self._s()), 'islam_pvgo_lm_step')
        return st

    def lm_run(self):
        st = LMState()
        self._enter()
        _lib.check(self.L.islam_pvgo_lm_run(self._h, C.byref(st), self._s()), 'islam_pvgo_lm_run')
        return st

    def profile_try(self):
        """One try with CUDA events between its phases -> dict of milliseconds."""
        ms = (C.c_float * 5)()
        self._enter()
        _lib.check(self.L.islam_pvgo_profile_try(self._h, C.byref(ms), self._s()), 'islam_pvgo_profile_try')
        return dict(linearize=ms[0], factor=ms[1], backsolve=ms[2], trial=ms[3], total=ms[4])

    def lm_try_async(self):
        _lib.check(self.L.islam_pvgo_lm_try(self._h, self._s()), 'islam_pvgo_lm_try')

    def lm_state(self):
        st = LMState()
        _lib.check(self.L.islam_pvgo_get_lm_state(self._h, C.byref(st), self._s()), 'islam_pvgo_get_lm_state')
        return st

    # ------------------------------------------------------------------------------------------------ outputs
    def vo_loss(self, vo_motions, with_grad=False):
        P = _f32(vo_motions, self.device, (self.E, 7))
        tl, rl = self._new(self.E), self._new(self.E)
        gt = self._new(self.E, 6) if with_grad else None
        gr = self._new(self.E, 6) if with_grad else None
        self._enter()
        _lib.check(self.L.islam_pvgo_vo_loss(self._h, _ptr(P), _ptr(tl), _ptr(rl), _ptr(gt), _ptr(gr), self._s()),
                   'islam_pvgo_vo_loss')
        self._exit()
        P.record_stream(self.stream)
        return (tl, rl, gt, gr) if with_grad else (tl, rl)

    def imu_loss(self, imu_drots=None, imu_dvels=None, with_grad=False):
        """pvgo.py:95-111 for the given measurements (None: the ones staged by set_problem)."""
        dr = _f32(imu_drots, self.device, (self.M, 4)) if imu_drots is not None else None
        dv = _f32(imu_dvels, self.device, (self.M, 3)) if imu_dvels is not None else None
        tl, rl = self._new(self.M), self._new(self.M)
        gr = self._new(self.M, 3) if with_grad else None
        gv = self._new(self.M, 3) if with_grad else None
        self._enter()
        _lib.check(self.L.islam_pvgo_imu_loss(self._h, _ptr(dr), _ptr(dv), _ptr(tl), _ptr(rl), _ptr(gr), _ptr(gv),
                                              self._s()), 'islam_pvgo_imu_loss')
        self._exit()
        for t in (dr, dv):
            if t is not None:
                t.record_stream(self.stream)
        return (tl, rl, gr, gv) if with_grad else (tl, rl)

    def align(self, target):
        t = _f32(target, self.device, (7,))
        n, v = self._new(self.N, 7), self._new(self.N, 3)
        self._enter()
        _lib.check(self.L.islam_pvgo_align(self._h, _ptr(t), _ptr(n), _ptr(v), self._s()), 'islam_pvgo_align')
        self._exit()
        t.record_stream(self.stream)
        return n, v
