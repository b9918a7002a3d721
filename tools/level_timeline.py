"""Developer probe: per-front globaltimer stamps of one factorisation (libislam_dbg.so, `python -m islam_b200.build
--phase-clocks`): for every level the wall time from the first CTA passing the grid dependency to the last CTA ending,
the spread of the per-front durations, and how many fronts shared an SM."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from islam_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), os.environ.get('ISLAM_DBG_LIB', 'libislam_dbg.so'))
import numpy as np, torch
from islam_b200 import synth
from islam_b200.solver import PVGOSolver
import mf_emul
g = synth.config2()
plan = mf_emul.get_plan(g.N, g.links)
s = PVGOSolver(g.N, g.links)
s.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight)
s.set_state(g.init_nodes, g.init_vels)
s.linearize()
L = C.CDLL(_lib.LIB_PATH)
for _ in range(3):
    s.solve(1.0001)
buf = (C.c_ulonglong * (4 * 8192))()
L.islam_debug_front_times(buf)
t = np.array(buf[:], dtype=np.uint64).reshape(4, 8192).astype(np.int64)
lv = plan['level']
t0 = t[1][:plan['F']].min()
prev_end = None
for l in range(plan['n_levels']):
    fs = np.where(lv == l)[0]
    fs = fs[fs < 8192]
    start, dep, end, sm = t[0][fs], t[1][fs], t[2][fs], t[3][fs]
    dur = (end - dep) / 1e3
    _, cnt = np.unique(sm, return_counts=True)
    gap = (dep.min() - prev_end) / 1e3 if prev_end is not None else 0.0
    print(f'level {l:2d} fronts {len(fs):4d}  level wall {(end.max() - dep.min()) / 1e3:7.1f} us  gap after previous level {gap:6.1f} us  '
          f'front dur min/med/max {dur.min():6.1f}/{np.median(dur):6.1f}/{dur.max():6.1f} us  '
          f'CTA resident before dep {np.median(dep - start) / 1e3:6.1f} us  max fronts per SM {cnt.max()}  SMs used {len(cnt)}')
    prev_end = end.max()
print('total', (t[2][:plan["F"]].max() - t0) / 1e3, 'us')
# back-substitution: CTA start / dependency (or parent counter) passed / end
L.islam_debug_bs_times(buf)
t = np.array(buf[:], dtype=np.uint64).reshape(4, 8192).astype(np.int64)
prev_end = None
print('back-substitution (root level first)')
for l in range(plan['n_levels'] - 1, -1, -1):
    fs = np.where(lv == l)[0]
    fs = fs[fs < 8192]
    start, dep, end = t[0][fs], t[1][fs], t[2][fs]
    dur = (end - dep) / 1e3
    gap = (dep.min() - prev_end) / 1e3 if prev_end is not None else 0.0
    print(f'level {l:2d} fronts {len(fs):4d}  level wall {(end.max() - dep.min()) / 1e3:7.1f} us  gap after parents {gap:6.1f} us  '
          f'front dur min/med/max {dur.min():6.1f}/{np.median(dur):6.1f}/{dur.max():6.1f} us  resident before go {np.median(dep - start) / 1e3:6.1f} us')
    prev_end = end.max()
print('total', (t[2][:plan['F']].max() - t[1][:plan['F']].min()) / 1e3, 'us')
