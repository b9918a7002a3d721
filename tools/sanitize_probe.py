"""Small end-to-end probe for compute-sanitizer (memcheck / racecheck / synccheck): a dense-root graph (potrf with inverse,
DMMA panel solve, TMA-staged update, multi-CTA back-substitution), a band graph through the multifrontal kernels, and the
one-launch small-window path with speculative retries."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from islam_b200 import synth
from islam_b200.solver import PVGOSolver
from islam_b200.pvgo import run_pvgo

what = sys.argv[1] if len(sys.argv) > 1 else 'all'
if what in ('all', 'root'):
    g = synth.config4(N=700, n_lc=45, min_gap=40)
    s = PVGOSolver(g.N, g.links)
    s.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight)
    s.set_state(g.init_nodes, g.init_vels)
    s.lm_reset(radius=g.radius, max_steps=2, use_scheduler=0)
    st = s.lm_run()
    print('dense root: root_pivots', s.dims.root_pivots, 'loss', st.loss, 'info', st.info, flush=True)
if what in ('all', 'band'):
    g = synth.config2(N=300, band=8)
    s = PVGOSolver(g.N, g.links)
    s.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight)
    s.set_state(g.init_nodes, g.init_vels)
    s.lm_reset(radius=g.radius, max_steps=2, use_scheduler=0)
    st = s.lm_run()
    print('band graph: loss', st.loss, 'info', st.info, flush=True)
if what in ('all', 'small'):
    g = synth.window()
    t = torch.as_tensor
    out = run_pvgo(t(g.init_nodes), t(g.init_vels), t(g.vo_motions), t(g.links), t(g.dts), t(g.imu_drots), t(g.imu_dtrans),
                   t(g.imu_dvels), radius=g.radius, loss_weight=g.loss_weight)
    st = run_pvgo.last_state
    print('small window: steps', st.steps_done, 'tries', st.tries_total, 'loss', st.loss, flush=True)
torch.cuda.synchronize()
print('probe done')
