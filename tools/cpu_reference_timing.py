"""SURVEY.md 8d 'CPU reference timing': (i) the literal-semantics dense oracle (PyPose's algorithm: dense J (R x 10N), dense
block-diagonal W (R x R), A = J^T W J, dense Cholesky) on C1 and on growing N to show the O(N^3) trend — the reference
itself cannot run C2 (W alone is 324 GB); (ii) the sparse CPU twin (same normal equations, block-sparse assembly, banded /
sparse-LU solve) on C2-C4.  Host cores of whatever box runs it; one LM iteration each (median of a few)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from islam_b200 import synth
from oracle import pvgo_oracle as po


def t_step(lm, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        lm.step()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


print(f'host cores: {os.cpu_count()}')
for name, g, reps in (('C1 (N=100)', synth.config1(), 3), ('band-8 N=200', synth.config2(N=200, band=8), 2),
                      ('band-8 N=400', synth.config2(N=400, band=8), 1)):
    lm = po.DenseLM(g, np.float32)
    rows = 6 * len(g.links) + 9 * (g.N - 1)
    print(f'dense oracle (PyPose-literal, float32)  {name:14s} rows {rows:6d} x cols {10 * g.N:5d}: {t_step(lm, reps) * 1e3:10.1f} ms per LM iteration')
for name, g, reps, kw in (('C2', synth.config2(), 3, {}), ('C3', synth.config3(), 3, {}),
                          ('C4 (50 000 poses, 2 000 closures)', synth.config4(), 1, dict(solver='splu'))):
    lm = po.SparseLM(g, np.float64, **kw)
    print(f'sparse CPU twin (float64)               {name:34s}: {t_step(lm, reps) * 1e3:10.1f} ms per LM iteration')
