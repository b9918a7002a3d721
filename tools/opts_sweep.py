"""Developer probe: factor / back-substitution time of one C2 LM try for different symbolic options (leaf_max, pivot_max)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from islam_b200 import synth
from islam_b200.solver import PVGOSolver
g = synth.config2()
for leaf_max, pivot_max in ((8, 8), (6, 8), (10, 8), (12, 8), (16, 8), (12, 12), (16, 12), (24, 12), (8, 6)):
    try:
        s = PVGOSolver(g.N, g.links, leaf_max=leaf_max, pivot_max=pivot_max)
    except Exception as e:
        print(leaf_max, pivot_max, 'create failed', e); continue
    s.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight)
    s.set_state(g.init_nodes, g.init_vels)
    s.lm_reset(radius=g.radius, max_steps=10, use_scheduler=0)
    st = s.lm_run()
    s.set_state(g.init_nodes, g.init_vels)
    s.lm_reset(radius=g.radius, max_steps=10, use_scheduler=0)
    ph = [s.profile_try() for _ in range(10)]
    m = {k: float(np.mean([p[k] for p in ph[2:]])) for k in ph[0]}
    print(f'leaf_max {leaf_max:2d} pivot_max {pivot_max:2d}: fronts {s.dims.F:5d} levels {s.dims.levels:2d} flops {s.dims.factor_flops/1e6:7.1f} MF  '
          f'lin {m["linearize"]:.3f} factor {m["factor"]:.3f} backsolve {m["backsolve"]:.3f} trial {m["trial"]:.3f} total {m["total"]:.3f} ms  loss {st.loss:.6f}')
    del s
