"""Developer probe: per-phase clock64() stamps of k_factor3 (block 0 of the level whose grid size is argv[1]) from a
-DISLAM_PHASE_CLOCKS build (`python -m islam_b200.build --phase-clocks`)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from islam_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), os.environ.get('ISLAM_DBG_LIB', 'libislam_dbg.so'))
import numpy as np, torch
from islam_b200 import synth
from islam_b200.solver import PVGOSolver
g = synth.config2()
s = PVGOSolver(g.N, g.links)
s.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight)
s.set_state(g.init_nodes, g.init_vels)
s.linearize()
L = C.CDLL(_lib.LIB_PATH)
for grid in [int(a) for a in sys.argv[1:]] or [1]:
    L.islam_debug_phase_grid(grid)
    for _ in range(3):
        s.solve(1.0001)
    buf = (C.c_longlong * 64)()
    L.islam_debug_phase_clocks(buf)
    c = np.array(buf[:], dtype=np.int64)
    print('level with grid', grid, 'block 0:')
    names = {0: 'start', 1: 'zeroed + previous level done', 2: 'original entries', 3: 'children added', 4: 'panel done',
             5: 'L written', 6: 'U done'}
    t0 = c[0]
    for k in (0, 1, 2, 3):
        print(' ', names[k], c[k] - t0)
    for jb in range(8):
        if c[10 + 3 * jb] > t0:
            print('  jb', jb, 'start', c[10 + 3 * jb] - t0, 'trsm', c[11 + 3 * jb] - t0, 'update', c[12 + 3 * jb] - t0)
    for k in (4, 5, 6):
        print(' ', names[k], c[k] - t0)
    print('  front4 jb 2 detail: loads done', c[8] - t0, 'diag done', c[9] - t0, 'global stores issued', c[5] - t0, 'look-ahead init loaded', c[6] - t0, 'look-ahead fma done', c[7] - t0)
    if c[30] > t0:          # front4: panel / Schur warps
        for jb in range(7):
            if c[30 + 3 * jb] > t0:
                print('  panel jb', jb, 'chain flag seen', c[30 + 3 * jb] - t0, 'row solve done', c[31 + 3 * jb] - t0, 'trailing + store done', c[32 + 3 * jb] - t0)
        for jb in range(6):
            if c[48 + 2 * jb] > t0:
                print('  schur jb', jb, 'panel flag seen', c[48 + 2 * jb] - t0, 'rank-9 done', c[49 + 2 * jb] - t0)
        print('  schur applied', c[60] - t0, 'panel warps done', c[61] - t0)
