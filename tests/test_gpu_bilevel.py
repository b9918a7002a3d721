"""The imperative-learning loop of /root/reference/train.py:162-299 on the B200 back-end with a synthetic learnable front-end
(BASELINE config 5's structure: VO forward -> chain poses -> IMU pre-integration -> run_pvgo -> one-step backprop -> optimiser).
The stand-in front-end has a systematic se3 bias; its rotational part is exposed by the IMU-fused PVGO estimate and must be
trained away through the one-step gradient of vo_loss."""
import numpy as np
import pytest
import torch

import islam_b200.pypose_compat as pp
from islam_b200 import synth
from islam_b200.imu_integrator import IMUModule
from islam_b200.pvgo import run_pvgo
from islam_b200.transformation import motion2pose_pypose, pose2motion_pypose
from oracle import lie

pytestmark = pytest.mark.gpu


def test_imperative_loop_trains_the_front_end():
    dev = 'cuda:0'
    batch, n_win = 8, 5                                          # run_kitti.sh:8 batch_size
    N = batch * n_win + 1
    gt, gv, _, _ = synth.ground_truth(N)
    imu = synth.raw_imu(N, sig_a=0.0, sig_g=0.0)
    imu_module = IMUModule(imu['accels'], imu['gyros'], imu['dts'], init=imu['init'], gravity=imu['gravity'],
                           rgb2imu_sync=imu['rgb2imu_sync'], device=dev, denoise_accel=False, denoise_gyro=False)
    rel = lie.se3_mul(lie.se3_inv(gt[:-1].astype(np.float64)), gt[1:].astype(np.float64))
    bias_true = np.array([0.03, -0.02, 0.01, 0.004, -0.003, 0.005])
    meas = torch.tensor(lie.se3_mul(rel, lie.se3_exp(bias_true[None])), dtype=torch.float32, device=dev)

    class FrontEnd(torch.nn.Module):                             # stands in for TartanVO: raw estimate * learnable correction
        def __init__(self):
            super().__init__()
            self.b = torch.nn.Parameter(torch.zeros(6, device=dev))

        def forward(self, idx):
            corr = pp.se3(self.b.unsqueeze(0)).Exp()
            return pp.SE3(meas[idx]) @ corr                      # LieTensor with autograd history (train.py:215)

    vo = FrontEnd()
    opt = torch.optim.Adam(vo.parameters(), lr=4e-3)
    loss_weight = (1.0, 0.1, 10.0, 0.1)
    history, bias_err = [], []
    for epoch in range(10):                                      # "imperative iterations"
        opt.zero_grad()
        init_state = dict(pos=gt[0, :3], rot=gt[0, 3:], vel=gv[0])
        total = 0.0
        for w in range(n_win):
            st, end = w * batch, (w + 1) * batch
            idx = torch.arange(st, end, device=dev)
            motions = vo(idx)
            T0 = pp.SE3(np.concatenate([init_state['pos'], init_state['rot']]).astype(np.float32))
            poses = motion2pose_pypose(motions, T0)              # train.py:220
            assert poses.shape == (batch + 1, 7)
            imu_trans, imu_rots, _, imu_vels = imu_module.integrate(st, end, init_state, motion_mode=False)   # :236
            imu_poses = pp.SE3(torch.cat((imu_trans, imu_rots.tensor()), axis=1))
            _ = pose2motion_pypose(imu_poses)                    # :240
            imu_dtrans, imu_drots, _, imu_dvels = imu_module.integrate(st, end, init_state, motion_mode=True)  # :244
            links = torch.stack([torch.arange(batch), torch.arange(1, batch + 1)], 1)
            dts = torch.full((batch,), 0.1)
            trans_loss, rot_loss, pgo_poses, pgo_vels, covs = run_pvgo(
                imu_poses, imu_vels, motions, links, dts, imu_drots, imu_dtrans, imu_dvels,
                device=dev, radius=1e4, loss_weight=loss_weight, target='vo')                                   # :256-263
            loss_bp = torch.cat((1.0 * rot_loss, 0.1 * trans_loss))                                           # :280
            assert loss_bp.requires_grad
            loss_bp.backward(torch.ones_like(loss_bp))                                                        # :283
            total += float(loss_bp.detach().sum())
            p = pgo_poses.numpy()
            init_state = dict(pos=p[-1][:3], rot=p[-1][3:] / np.linalg.norm(p[-1][3:]), vel=pgo_vels[-1].numpy())  # :297-299
        opt.step()                                               # once per trajectory (train.py:175)
        history.append(total)
        bias_err.append(float(np.linalg.norm((vo.b.detach().cpu().numpy() + bias_true)[3:])))
    assert history[-1] < 0.5 * history[0], history              # the outer loss falls
    # with loss_weight (1, 0.1, 10, 0.1) the IMU pins rotation (info 100 vs 1) but hardly translation (0.01 vs 1): the
    # rotational part of the bias is what PVGO exposes and what imperative training removes
    rot0 = float(np.linalg.norm(bias_true[3:]))
    print('outer loss', history, 'rotation-bias error', bias_err, 'initial', rot0)
    assert bias_err[-1] < 0.5 * rot0, (bias_err, rot0)
