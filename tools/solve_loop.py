"""Developer driver for ncu: linearise C2 once, then a few factor + solve passes (10 k_factor3 + 10 k_backsolve3 launches each)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from islam_b200 import synth
from islam_b200.solver import PVGOSolver
g = synth.config2()
s = PVGOSolver(g.N, g.links)
s.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight)
s.set_state(g.init_nodes, g.init_vels)
s.linearize()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    s.solve(1.0001)
torch.cuda.synchronize()
