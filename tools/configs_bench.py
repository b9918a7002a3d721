"""Measurement of every BASELINE.json config on one B200 (CUDA events on the solver's stream, after warm-up):
C1 100-pose chain (5 LM iterations), C2 (the bench.py workload), C3 KITTI-00 length incl. GPU IMU pre-integration,
C4 50 000 poses + 2 000 loop closures (dense root), C5's back-end share: run_pvgo on 9-pose windows (run_kitti.sh:8), host in/out."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from islam_b200 import synth
from islam_b200.solver import PVGOSolver
from islam_b200.pvgo import run_pvgo
from islam_b200.imu_integrator import IMUModule


def timed(s, g, steps, reps):
    out = []
    for _ in range(reps + 2):
        s.set_state(g.init_nodes, g.init_vels)
        s.lm_reset(radius=g.radius, max_steps=steps, use_scheduler=0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s.stream):
            e0.record(); st = s.lm_run(); e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return float(np.median(out[2:])), st


def solver(g):
    t0 = time.perf_counter()
    s = PVGOSolver(g.N, g.links)
    s.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight)
    torch.cuda.synchronize()
    return s, time.perf_counter() - t0


for name, g, steps, reps in (('C1 100 poses / 300 factors, 5 iterations', synth.config1(), 5, 20),
                             ('C2 5 000 poses / 49 962 factors, 10 iterations', synth.config2(), 10, 10)):
    s, tc = solver(g)
    ms, st = timed(s, g, steps, reps)
    print(f'{name}: {ms:8.3f} ms per solve = {ms / st.tries_total:7.4f} ms per LM iteration ({1e3 * st.tries_total / ms:8.1f} it/s), '
          f'{s.dims.F} fronts / {s.dims.levels} levels, symbolic analysis + upload {tc * 1e3:6.1f} ms (once per graph structure)')

# C3: raw 100 Hz IMU -> pre-integration on the GPU (both modes, train.py:236,244) -> PVGO
N = 4541
g = synth.config3(N=N)
imu = synth.raw_imu(N)
m = IMUModule(imu['accels'], imu['gyros'], imu['dts'], init=imu['init'], gravity=imu['gravity'], rgb2imu_sync=imu['rgb2imu_sync'],
              device='cuda:0', denoise_accel=False, denoise_gyro=False)
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pos, rot, _, vel = m.integrate(0, N - 1, imu['init'], motion_mode=False)
    dtrans, drots, _, dvels = m.integrate(0, N - 1, imu['init'], motion_mode=True)
    torch.cuda.synchronize(); t_imu = time.perf_counter() - t0
g.init_nodes = np.concatenate([pos.numpy(), torch.as_tensor(rot).numpy()], 1).astype(np.float32)
g.init_vels = vel.numpy().astype(np.float32)
g.imu_drots, g.imu_dtrans, g.imu_dvels = torch.as_tensor(drots).numpy(), dtrans.numpy(), dvels.numpy()
s, tc = solver(g)
ms, st = timed(s, g, 10, 10)
print(f'C3 4 541 poses (KITTI-00 length), 45 400 IMU samples: pre-integration (world + motion mode, results on the host) {t_imu * 1e3:7.3f} ms; '
      f'PVGO {ms:8.3f} ms per solve = {ms / st.tries_total:7.4f} ms per LM iteration, {s.dims.F} fronts / {s.dims.levels} levels')

g = synth.config4()
s, tc = solver(g)
ms, st = timed(s, g, 3, 1)
print(f'C4 50 000 poses / 51 999 edges / 2 000 loop closures: {ms / st.tries_total:8.1f} ms per LM iteration, dense root of '
      f'{s.dims.root_pivots} poses, {s.dims.F} fronts / {s.dims.levels} levels, symbolic analysis + upload {tc:5.2f} s')
del s

g = synth.window()
t = lambda a: torch.as_tensor(a).pin_memory()
args = [t(g.init_nodes), t(g.init_vels), t(g.vo_motions), torch.as_tensor(g.links), t(g.dts), t(g.imu_drots), t(g.imu_dtrans), t(g.imu_dvels)]
for _ in range(5):
    run_pvgo(*args, device='cuda:0', radius=g.radius, loss_weight=g.loss_weight)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50):
    out = run_pvgo(*args, device='cuda:0', radius=g.radius, loss_weight=g.loss_weight)
torch.cuda.synchronize()
print(f'C5 back-end share: run_pvgo on a 9-pose window (host tensors in and out, StopOnPlateau): {(time.perf_counter() - t0) / 50 * 1e3:7.3f} ms per call, '
      f'{run_pvgo.last_state.steps_done} LM iterations')
