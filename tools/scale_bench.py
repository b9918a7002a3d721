"""Measurement of SURVEY 8f rank 4 (islam_scale_from_disp_flow) on its real shape: TartanVO's quarter-resolution grid
160 x 112 (640 x 448 / 4, TartanVO.py:121), batch 8 (run_kitti.sh:8), plus a large batch that is actually HBM-sized.
CUDA events around 200 launches after warm-up; algorithmic bytes = 12 read + 6 written per pixel; the float64 NumPy oracle
timed beside it on one host core (per-sample loop, as the reference's Python loop at TartanVO.py:159-171)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from islam_b200 import dense_ba
from oracle import dense_ba_oracle as dbo

peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs'] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')) else 6650.0
for B in (8, 4096):
    H, W = 112, 160
    g = torch.Generator(device='cuda').manual_seed(0)
    disp = 1 + 10 * torch.rand(B, H, W, device='cuda', generator=g)
    flow = 5 * torch.randn(B, 2, H, W, device='cuda', generator=g)
    mo = torch.tensor([[0.1, 0.0, 0.99, 0, 0, 0, 1.0]], device='cuda').expand(B, 7).contiguous()
    intr = torch.tensor([[80.0, 80.0, 79.5, 55.5]], device='cuda').expand(B, 4).contiguous()
    bl = torch.full((B,), 0.5, device='cuda')
    for _ in range(5):
        out = dense_ba.scale_from_disp_flow_batch(disp, flow, mo, intr, bl)
    torch.cuda.synchronize()
    n = 200 if B == 8 else 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = dense_ba.scale_from_disp_flow_batch(disp, flow, mo, intr, bl)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    bytes_ = 18.0 * B * H * W
    line = f'batch {B} x {H} x {W}: {us:9.1f} us per call (incl. output allocation + ctypes), {B * H * W / us:10.1f} Mpixel/s, ' \
           f'{bytes_ / us / 1e3:8.1f} GB/s algorithmic = {bytes_ / us / 1e3 / peak:6.3f} of the measured HBM peak ({peak:.0f} GB/s)'
    if B == 8:
        d, f = disp.cpu().numpy(), flow.cpu().numpy()
        t0 = time.perf_counter()
        for i in range(B):
            dbo.scale_from_disp_flow(d[i], f[i], mo[i].cpu().numpy(), 80.0, 80.0, 79.5, 55.5, 0.5)
        line += f'; NumPy oracle, per-sample loop on one host core: {(time.perf_counter() - t0) * 1e6:9.1f} us'
    print(line)
