"""Drop-in mirror of /root/reference/imu_integrator.py (IMUModule) on top of the fused pre-integration kernels.

    IMUModule(accels, gyros, dts, accel_bias, gyro_bias, init, gravity, rgb2imu_sync, device, denoise_model_name,
              denoise_accel, denoise_gyro, use_est_cov).integrate(st, end, init, motion_mode)
        -> (poses (K,3) cpu, rots SO3 (K,4) cpu, covs [], vels (K,3) cpu)           imu_integrator.py:31-164

The reference runs one PyPose IMUPreintegrator call (plus three .cpu() syncs) per camera frame in a Python loop;
here the whole [st, end) window is three kernel launches (csrc/imu.cu).  The learned denoiser
(Network/IMUDenoiseNet.py) is a front-end network outside this path: passing denoise_model_name raises.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import IslamError


def prase_init(init=None, motion_mode=False, device='cuda:0'):
    """imu_integrator.py:11-28 (name kept as spelled there) -> (pos(3), rot(4 xyzw), vel(3)) float32 on `device`."""
    dtype = torch.get_default_dtype()
    if init is not None:
        rot = torch.as_tensor(np.asarray(_to_np(init['rot'])), dtype=dtype).reshape(4)
        if motion_mode:
            pos, vel = torch.zeros(3, dtype=dtype), torch.zeros(3, dtype=dtype)
        else:
            pos = torch.as_tensor(np.asarray(_to_np(init['pos'])), dtype=dtype).reshape(3)
            vel = torch.as_tensor(np.asarray(_to_np(init['vel'])), dtype=dtype).reshape(3)
    else:
        pos, vel = torch.zeros(3, dtype=dtype), torch.zeros(3, dtype=dtype)
        rot = torch.tensor([0., 0., 0., 1.], dtype=dtype)
    return pos.to(device), _wrap_so3(rot.to(device)), vel.to(device)


def _plain(t):
    return t.as_subclass(torch.Tensor) if type(t) is not torch.Tensor else t


def _to_np(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return x


def _wrap_so3(t):
    try:
        from . import pypose_compat as pp
        return pp.SO3(t)
    except Exception:      # pragma: no cover
        return t


class IMUModule:
    def __init__(self, accels, gyros, dts, accel_bias=torch.zeros(3), gyro_bias=torch.zeros(3), init=None,
                 gravity=9.81007, rgb2imu_sync=None, device='cuda:0', denoise_model_name=None, denoise_accel=True,
                 denoise_gyro=True, use_est_cov=False):
        if not torch.cuda.is_available():
            raise IslamError('IMUModule needs a CUDA device: there is no CPU fallback')
        self.device = torch.device(device)
        self.L = _lib.lib()
        self.last_frame_dt = 0.1
        self.rgb2imu_sync = list(range(len(accels))) if rgb2imu_sync is None else rgb2imu_sync       # :38-41
        dtype = torch.float32
        self.accels = torch.as_tensor(np.asarray(_to_np(accels)), dtype=dtype).to(self.device).contiguous()
        self.gyros = torch.as_tensor(np.asarray(_to_np(gyros)), dtype=dtype).to(self.device).contiguous()
        self.dts = torch.as_tensor(np.asarray(_to_np(dts)), dtype=dtype).reshape(-1).to(self.device).contiguous()
        self.denoise_accel, self.denoise_gyro = denoise_accel, denoise_gyro
        self.use_denoise_model = denoise_model_name is not None and denoise_model_name != '' and \
            (denoise_accel or denoise_gyro)                                                          # :50
        if self.use_denoise_model:
            raise NotImplementedError('the CNN-GRU IMU denoiser (Network/IMUDenoiseNet.py) is a front-end model '
                                      'outside the B200 back-end path')
        self.optm_bias = not self.use_denoise_model and (denoise_accel or denoise_gyro)              # :51
        self.gravity = float(gravity)
        self.accel_bias = torch.as_tensor(np.asarray(_to_np(accel_bias)), dtype=dtype).to(self.device)
        self.gyro_bias = torch.as_tensor(np.asarray(_to_np(gyro_bias)), dtype=dtype).to(self.device)
        self._sync = torch.as_tensor(np.asarray(self.rgb2imu_sync), dtype=torch.int32).to(self.device)
        self.stream = torch.cuda.Stream(device=self.device)

    def integrate(self, st, end, init=None, motion_mode=False):
        """motion_mode False: pos/rot/vel in the world frame, chained from `init` (K = end-st+1 rows, init first).
        motion_mode True: rot = R_t^-1 R_{t+1}, vel = world-frame delta-v, pos = world-frame displacement caused by
        acceleration only (K = end-st rows).   imu_integrator.py:69-164."""
        init_pos, init_rot, init_vel = prase_init(init, motion_mode, self.device)
        K = int(end) - int(st)
        b0 = int(self.rgb2imu_sync[st])
        b1 = int(self.rgb2imu_sync[end]) + 1                                                         # :91-92
        S = b1 - b0
        acc, gyr, dts = self.accels[b0:b1], self.gyros[b0:b1], self.dts[b0:b1]
        if self.optm_bias:                                                                           # :101-105
            if self.denoise_accel:
                acc = acc - self.accel_bias.view(1, 3)
            if self.denoise_gyro:
                gyr = gyr - self.gyro_bias.view(1, 3)
        acc, gyr, dts = acc.contiguous(), gyr.contiguous(), dts.contiguous()
        off = (self._sync[st:end + 1] - b0).contiguous()
        init10 = torch.cat([init_pos, _plain(init_rot), init_vel]).to(torch.float32).contiguous()
        pos = torch.empty(K, 3, device=self.device)
        rot = torch.empty(K, 4, device=self.device)
        vel = torch.empty(K, 3, device=self.device)
        ws = torch.empty(int(self.L.islam_imu_workspace_bytes(S, K)), dtype=torch.uint8, device=self.device)
        p = lambda t: C.c_void_p(t.data_ptr())
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        _lib.check(self.L.islam_imu_preintegrate(p(acc), p(gyr), p(dts), S, p(off), K, p(init10), self.gravity,
                                                 1 if motion_mode else 0, p(pos), p(rot), p(vel), p(ws),
                                                 C.c_void_p(self.stream.cuda_stream)), 'islam_imu_preintegrate')
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        for t in (acc, gyr, dts, off, init10, ws):
            t.record_stream(self.stream)
        if not motion_mode:                                                                          # :86-89
            pos = torch.cat([init_pos.view(1, 3).float(), pos])
            rot = torch.cat([_plain(init_rot).view(1, 4).float(), rot])
            vel = torch.cat([init_vel.view(1, 3).float(), vel])
        # poses.cpu(), rots.cpu(), vels.cpu() (imu_integrator.py:148-164) as ONE packed copy into pinned memory and one sync
        packed = torch.cat([pos, rot, vel], dim=1)
        host = torch.empty(packed.shape, dtype=packed.dtype, pin_memory=True)
        host.copy_(packed, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return host[:, 0:3].clone(), _wrap_so3(host[:, 3:7].clone()), [], host[:, 7:10].clone()
