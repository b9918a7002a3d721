"""pp.module.IMUPreintegrator as used at /root/reference/imu_integrator.py:55-56,146, on the fused kernel."""
import ctypes as C

import torch

from .. import _lib
from .._lib import IslamError


class IMUPreintegrator(torch.nn.Module):
    """forward(dt, gyro, acc, init_state) -> {'pos','rot','vel'} of shape (1, F, .) for every sample k = 1..F.
    Covariance propagation (prop_cov) is not produced: the reference discards it (imu_integrator.py:84,88,164)."""

    def __init__(self, pos=None, rot=None, vel=None, gravity=9.81007, prop_cov=True, reset=False, **kw):
        super().__init__()
        z3 = torch.zeros(3)
        self.register_buffer('pos', torch.as_tensor(pos if pos is not None else z3).detach().float().reshape(-1)[:3].clone(), persistent=False)
        r = rot if rot is not None else torch.tensor([0., 0., 0., 1.])
        self.register_buffer('rot', torch.as_tensor(r).detach().as_subclass(torch.Tensor).float().reshape(-1)[:4].clone(), persistent=False)
        self.register_buffer('vel', torch.as_tensor(vel if vel is not None else z3).detach().float().reshape(-1)[:3].clone(), persistent=False)
        self.gravity = float(gravity)

    def forward(self, dt, gyro, acc, rot=None, gyro_cov=None, acc_cov=None, init_state=None):
        from . import LieTensor, SO3_type
        if not torch.cuda.is_available():
            raise IslamError('IMUPreintegrator runs on the CUDA kernels only')
        dev = acc.device if acc.is_cuda else torch.device('cuda', torch.cuda.current_device())
        f32 = lambda t, w: t.detach().as_subclass(torch.Tensor).to(dev, torch.float32).reshape(-1, w).contiguous()
        a, g, d = f32(acc, 3), f32(gyro, 3), f32(dt, 1).reshape(-1)
        F = a.shape[0]
        if init_state is not None:
            ip, ir, iv = init_state['pos'], init_state['rot'], init_state['vel']
        else:
            ip, ir, iv = self.pos, self.rot, self.vel
        init = torch.cat([f32(ip, 3)[0], f32(ir, 4)[0], f32(iv, 3)[0]]).contiguous()
        off = torch.arange(F + 1, dtype=torch.int32, device=dev)      # one "frame" per sample => every k is returned
        pos, rotq, vel = (torch.empty(F, w, device=dev) for w in (3, 4, 3))
        L = _lib.lib()
        ws = torch.empty(int(L.islam_imu_workspace_bytes(F, F)), dtype=torch.uint8, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            _lib.check(L.islam_imu_preintegrate(p(a), p(g), p(d), F, p(off), F, p(init), self.gravity, 0, p(pos), p(rotq),
                                                p(vel), p(ws), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                       'islam_imu_preintegrate')
        back = lambda t: t.to(acc.device).unsqueeze(0)
        return {'pos': back(pos), 'rot': LieTensor(back(rotq), ltype=SO3_type), 'vel': back(vel), 'cov': None}
