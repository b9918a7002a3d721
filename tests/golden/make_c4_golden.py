"""Generates tests/golden/c4_oracle_steps.npz: the float64 CPU oracle (sparse LU twin, identical normal equations) on the
FULL BASELINE config 4 — 50 000 poses, 51 999 edges, 2 000 loop closures, synth.config4(), seed 0 — for 8 optimizer.step
calls: the loss / reject count after every step and every 25th pose + velocity of the aligned state (pvgo.py:195) after
step 3 and after step 8.
tests/test_gpu_pvgo.py::test_config4_full_size_against_oracle_fixture compares the GPU run with it: full-size parity of the
dense-root path that no in-test oracle run could afford (~80 s of CPU per LM iteration).

    python tests/golden/make_c4_golden.py          (about 12 minutes of CPU, ~6 GB)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from islam_b200 import synth                     # noqa: E402
from oracle import pvgo_oracle as po             # noqa: E402

STEPS, STRIDE, MID = 8, 25, 3

if __name__ == '__main__':
    g = synth.config4()
    lm = po.SparseLM(g, np.float64, solver='splu')
    for k in range(STEPS):
        t0 = time.time()
        lm.step()
        print('step', k, 'loss', lm.history[-1]['loss'], 'rejects', lm.history[-1]['rejects'], '%.0f s' % (time.time() - t0), flush=True)
        if k + 1 == MID:
            n_mid, v_mid = [a[::STRIDE].copy() for a in lm.aligned(g.init_nodes[0])]
    n, v = lm.aligned(g.init_nodes[0])
    np.savez_compressed(os.path.join(HERE, 'c4_oracle_steps.npz'), nodes=n[::STRIDE], vels=v[::STRIDE], nodes_mid=n_mid, vels_mid=v_mid,
                        mid=np.int32(MID),
                        losses=np.array([h['loss'] for h in lm.history]), rejects=np.array([h['rejects'] for h in lm.history]),
                        steps=np.int32(STEPS), stride=np.int32(STRIDE), N=np.int32(g.N), E=np.int32(g.E))
    print('C4 oracle fixture written')
