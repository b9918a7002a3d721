"""the NumPy restatement of scale_from_disp_flow against golden vectors produced by the reference function
itself (tests/golden/make_scale_golden.py), plus the closed-form property that pins the conventions."""
import os

import numpy as np

from oracle import dense_ba_oracle as dbo

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'scale_golden.npz'))


def case(k):
    kw = dict(depth=G[f'{k}_depth'] if bool(G[f'{k}_has_depth']) else None, mask=G[f'{k}_mask'] if bool(G[f'{k}_has_mask']) else None,
              disp_th=float(G[f'{k}_disp_th']))
    fx, fy, cx, cy = [float(x) for x in G[f'{k}_intr']]
    return (G[f'{k}_disp'], G[f'{k}_flow'], G[f'{k}_motion'], fx, fy, cx, cy, float(G[f'{k}_baseline'])), kw


def test_restatement_matches_the_reference_outputs():
    for k in range(int(G['n'])):
        args, kw = case(k)
        s, z, m, dm = dbo.scale_from_disp_flow(*args, **kw)
        assert np.array_equal(dm, G[f'{k}_ref_dmask']), k
        # float32 vs float64 can flip a pixel exactly on a mask boundary; the scenes have none
        assert np.array_equal(m, G[f'{k}_ref_mask']), k
        assert np.abs(z - G[f'{k}_ref_z']).max() <= 1e-5 * np.abs(G[f'{k}_ref_z']).max(), k
        assert abs(s - float(G[f'{k}_ref_s'][0])) <= 2e-4 * abs(s), (k, s, G[f'{k}_ref_s'])


def test_exact_flow_recovers_the_true_scale():
    for k in range(int(G['n'])):
        if str(G[f'{k}_kind']) != 'exact':
            continue
        args, kw = case(k)
        s, _, m, _ = dbo.scale_from_disp_flow(*args, **kw)
        assert m.sum() > 500
        assert abs(s - float(G[f'{k}_s_true'])) <= 1e-4 * s
