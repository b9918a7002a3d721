"""Measurement of the small-graph fast path (csrc/small.cuh) on the reference's shipped window size (9 poses, run_kitti.sh:8):
run_pvgo end to end (host tensors in and out) through the one-launch path and through the general multifrontal path, the
kernel alone (CUDA events), and run_pvgo_batch throughput."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from islam_b200 import synth
from islam_b200.pvgo import run_pvgo, run_pvgo_batch

g = synth.window()
t = lambda a: torch.as_tensor(a).pin_memory()
args = [t(g.init_nodes), t(g.init_vels), t(g.vo_motions), torch.as_tensor(g.links), t(g.dts), t(g.imu_drots), t(g.imu_dtrans), t(g.imu_dvels)]


def e2e(n=200):
    for _ in range(10):
        run_pvgo(*args, device='cuda:0', radius=g.radius, loss_weight=g.loss_weight)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        run_pvgo(*args, device='cuda:0', radius=g.radius, loss_weight=g.loss_weight)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, run_pvgo.last_state


ms, st = e2e()
print(f'run_pvgo, 9-pose window, one-launch path: {ms:7.4f} ms per call end to end ({st.steps_done} LM steps, {st.tries_total} tries)')
os.environ['ISLAM_SMALL_SPEC'] = '0'
ms1, st1 = e2e()
del os.environ['ISLAM_SMALL_SPEC']
print(f'run_pvgo, same, speculative retries off   : {ms1:7.4f} ms per call end to end ({st1.steps_done} LM steps, {st1.tries_total} tries)')
os.environ['ISLAM_NO_SMALL'] = '1'
ms2, st2 = e2e(50)
del os.environ['ISLAM_NO_SMALL']
print(f'run_pvgo, 9-pose window, general path   : {ms2:7.4f} ms per call end to end ({st2.steps_done} LM steps, {st2.tries_total} tries)')
# kernel alone: device-resident inputs, CUDA events
from islam_b200 import small
for B in (1, 148, 1184, 4736):
    gs = [synth.window(seed=s % 64) for s in range(B)]
    stack = lambda k: torch.as_tensor(np.stack([getattr(x, k) for x in gs])).cuda()
    a = dict(nodes0=stack('init_nodes'), vels0=stack('init_vels'), Z=stack('vo_motions'), drot=stack('imu_drots'),
             dtrans=stack('imu_dtrans'), dvel=stack('imu_dvels'), dt=stack('dts'))
    r = small.get_runner(9, g.links, 'cuda:0', B)
    for _ in range(3):
        r.run(a, g.loss_weight, radius=g.radius)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        states, *_ = r.run(a, g.loss_weight, radius=g.radius)
    e1.record(); torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / 10
    tries = sum(s.tries_total for s in states)
    print(f'batch of {B:5d} windows: {per:8.4f} ms per launch + copies = {per / B * 1e3:8.3f} us per window, {tries / (per * 1e-3):12.0f} LM tries/s')
