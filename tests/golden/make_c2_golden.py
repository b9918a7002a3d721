"""Generates tests/golden/c2_oracle_final.npz: the float64 CPU oracle's solution of BASELINE config C2
(5 000 poses / 49 962 factors, synth.config2(), seed 0) after exactly 10 optimizer.step calls, aligned to the first
initial pose (pvgo.py:195), plus the per-step losses.  bench.py compares the GPU result of EVERY run (any number of GPUs)
with this fixture and prints `lm.rel_pose_error_vs_oracle` — the parity gate of BASELINE.json's north_star (<= 1e-5
relative pose error after the same iteration count) carried by the bench line itself.

    python tests/golden/make_c2_golden.py          (about 15 s of CPU)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from islam_b200 import synth                     # noqa: E402
from oracle import pvgo_oracle as po             # noqa: E402

if __name__ == '__main__':
    g = synth.config2()
    lm = po.SparseLM(g, np.float64)
    lm.run(steps=10)
    n, v = lm.aligned(g.init_nodes[0])
    np.savez_compressed(os.path.join(HERE, 'c2_oracle_final.npz'), nodes=n, vels=v,
                        losses=np.array([h['loss'] for h in lm.history]), steps=np.int32(10), seed=np.int32(0),
                        N=np.int32(g.N), E=np.int32(g.E))
    print('C2 oracle: final loss', lm.history[-1]['loss'], 'steps', len(lm.history))
