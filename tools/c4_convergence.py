"""Developer probe: loss / rejects / damping of 60 LM steps on the full BASELINE config 4 (how far from a fixed point the
steps compared in tests/test_gpu_pvgo.py::test_config4_full_size_against_oracle_fixture are)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from islam_b200 import synth
from islam_b200.solver import PVGOSolver
g=synth.config4()
s=PVGOSolver(g.N,g.links); s.set_problem(g.vo_motions,g.imu_drots,g.imu_dtrans,g.imu_dvels,g.dts,g.loss_weight); s.set_state(g.init_nodes,g.init_vels)
s.lm_reset(radius=g.radius,max_steps=60,use_scheduler=0)
for k in range(60):
    st=s.lm_step()
    print(k+1, '%.6f'%st.loss, st.reject_count, st.tries_total, '%.3e'%st.damping, flush=True)
