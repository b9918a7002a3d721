"""Developer probe (GPU box): run a synthetic config through the LM driver, print per-step state, parity vs the oracle
and CUDA-event timings.  Not part of the product or the test-suite."""
import argparse, json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from islam_b200 import synth
from islam_b200.solver import PVGOSolver

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='C2')
ap.add_argument('--steps', type=int, default=10)
ap.add_argument('--oracle', action='store_true')
ap.add_argument('--stepwise', action='store_true')
ap.add_argument('--reps', type=int, default=5)
ap.add_argument('--leaf', type=int, default=0)
ap.add_argument('--pivot', type=int, default=0)
a = ap.parse_args()
g = {'C1': synth.config1, 'C2': synth.config2, 'C3': synth.config3, 'win9': synth.window,
     'C4s': lambda: synth.config4(N=5000, n_lc=20), 'C4': synth.config4}[a.config]()
t0 = time.time()
s = PVGOSolver(g.N, g.links, leaf_max=a.leaf, pivot_max=a.pivot)
d = s.dims
print('create %.3fs' % (time.time() - t0), dict(N=d.N, E=d.E, P=d.P, F=d.F, levels=d.levels, band=d.band, root=d.root_pivots,
      max_rows=d.max_rows, max_cols=d.max_cols, L_MB=d.L_doubles * 8 / 1e6, U_MB=d.U_doubles * 8 / 1e6, mflop=d.factor_flops / 1e6))
s.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight)
s.set_state(g.init_nodes, g.init_vels)
if a.stepwise:
    s.lm_reset(radius=g.radius, max_steps=a.steps, use_scheduler=0)
    for k in range(a.steps):
        st = s.lm_step()
        print(k, {k2: v for k2, v in st.as_dict().items() if k2 in ('loss', 'last', 'loss_trial', 'damping', 'quality', 'reject_count', 'tries_total', 'info', 'steps_done')})
else:
    for rep in range(a.reps):
        s.set_state(g.init_nodes, g.init_vels)
        s.lm_reset(radius=g.radius, max_steps=a.steps, use_scheduler=0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s.stream):
            e0.record()
            st = s.lm_run()
            e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print('rep', rep, 'ms total %.3f' % ms, 'per try %.3f' % (ms / max(1, st.tries_total)), 'steps', st.steps_done, 'tries', st.tries_total, 'loss', st.loss, 'info', st.info)
n, v = s.align(g.init_nodes[0])
if a.oracle:
    from oracle import pvgo_oracle as po
    ref = po.SparseLM(g, np.float64).run(steps=a.steps)
    for h in ref.history: print('oracle', h)
    rn, rv = ref.aligned(g.init_nodes[0])
    print('parity', po.rel_pose_error(n.cpu().numpy(), rn), 'vel', float(np.abs(v.cpu().numpy() - rv).max()))
