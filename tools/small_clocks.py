import os, sys, ctypes as C
sys.path.insert(0, '/root/repo')
from islam_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), 'libislam_dbg.so')
import numpy as np, torch
from islam_b200 import synth
from islam_b200.pvgo import run_pvgo
g = synth.window()
t = torch.as_tensor
a = [t(g.init_nodes), t(g.init_vels), t(g.vo_motions), t(g.links), t(g.dts), t(g.imu_drots), t(g.imu_dtrans), t(g.imu_dvels)]
for _ in range(3):
    run_pvgo(*a, radius=g.radius, loss_weight=g.loss_weight)
L = C.CDLL(_lib.LIB_PATH)
buf = (C.c_longlong * 16)()
L.islam_debug_small_clocks(buf)
c = np.array(buf[:])
names = ['try start', 'factors+zero (if lin)', 'assembly/damp', 'cholesky', 'backsolve', 'retract', 'trial+control']
for k in range(1, 7):
    print(names[k], c[k] - c[k - 1] if k != 1 else c[1] - c[0])
print('last try total', c[6] - c[0])
