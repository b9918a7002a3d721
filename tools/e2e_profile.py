"""Host-side profile of run_pvgo on C2 (where do the 0.7 ms between the device-timed solve and the end-to-end call go?)."""
import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from islam_b200 import synth
from islam_b200.pvgo import run_pvgo
g = synth.config2()
t = lambda a: torch.as_tensor(a).pin_memory()
args = [t(g.init_nodes), t(g.init_vels), t(g.vo_motions), torch.as_tensor(g.links), t(g.dts), t(g.imu_drots), t(g.imu_dtrans), t(g.imu_dvels)]
kw = dict(device='cuda:0', radius=g.radius, loss_weight=g.loss_weight, use_scheduler=False, max_steps=10)
for _ in range(5):
    run_pvgo(*args, **kw)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    run_pvgo(*args, **kw)
torch.cuda.synchronize()
print('ms per call', (time.perf_counter() - t0) / 20 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    run_pvgo(*args, **kw)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('cumulative').print_stats(35)
