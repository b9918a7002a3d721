"""Developer probe: per-phase clock64() stamps of k_factor_fast (block 0) from a -DISLAM_PHASE_CLOCKS build."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from islam_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), 'libislam_dbg.so')
import numpy as np, torch
from islam_b200 import synth
from islam_b200.solver import PVGOSolver
g = synth.config2()
s = PVGOSolver(g.N, g.links)
s.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight)
s.set_state(g.init_nodes, g.init_vels)
s.linearize()
L = C.CDLL(_lib.LIB_PATH)
L.islam_debug_phase_grid(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
for _ in range(3):
    s.solve(1.0001)
buf = (C.c_longlong * 64)()
L.islam_debug_phase_clocks(buf)
c = np.array(buf[:], dtype=np.int64)
print('level with grid', sys.argv[1:] , 'block 0:')
names = {0: 'start', 1: 'staged', 2: 'assembled', 3: 'panel done', 4: 'L written', 5: 'U done'}
t0 = c[0]
for k in (0, 1, 2):
    print(names[k], c[k] - t0)
for jb in range(8):
    print('jb', jb, 'chol', c[10 + 3 * jb] - t0, 'trsm', c[11 + 3 * jb] - t0, 'update', c[12 + 3 * jb] - t0)
for k in (3, 4, 5):
    print(names[k], c[k] - t0)

print('A iter0: issue done', c[41]-c[40], 'sum+store done', c[42]-c[40], 'iter1 start', c[43]-c[40])
