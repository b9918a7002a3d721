"""The small-graph fast path (csrc/small.cuh: the whole run_pvgo of a window in one launch) through the C ABI, against the
oracle and against the general multifrontal path: the shipped window size (9 poses, run_kitti.sh:8) with the plateau
scheduler and its 16-reject step, a badly perturbed window (rejected tries), the largest supported window with a loop
closure (full bandwidth), the autograd of vo_loss, and batches of windows (one CTA each)."""
import os

import numpy as np
import pytest
import torch

from islam_b200 import synth
from islam_b200 import pvgo as ipvgo
from islam_b200.pvgo import run_pvgo, run_pvgo_batch
from oracle import lie, pvgo_oracle as po

pytestmark = pytest.mark.gpu
_t = torch.as_tensor


def _args(g):
    return (_t(g.init_nodes), _t(g.init_vels), _t(g.vo_motions), _t(g.links), _t(g.dts), _t(g.imu_drots), _t(g.imu_dtrans),
            _t(g.imu_dvels))


def _general(g, **kw):
    os.environ['ISLAM_NO_SMALL'] = '1'
    try:
        out = run_pvgo(*_args(g), radius=g.radius, loss_weight=g.loss_weight, **kw)
        return out, run_pvgo.last_state
    finally:
        del os.environ['ISLAM_NO_SMALL']


def test_window_matches_oracle_and_general_path():
    g = synth.window()
    ref = po.SparseLM(g, np.float64).run()                      # StopOnPlateau(10, 3, 1e-3): step 2 burns its 16 rejects
    tl, rl, nodes, vels, covs = run_pvgo(*_args(g), radius=g.radius, loss_weight=g.loss_weight)
    st = run_pvgo.last_state
    assert st.steps_done == len(ref.history) and st.reject_count == ref.history[-1]['rejects']
    assert abs(st.loss - ref.history[-1]['loss']) <= 1e-4 * abs(ref.history[-1]['loss'])
    rn, rv = ref.aligned(g.init_nodes[0])
    assert po.rel_pose_error(np.asarray(nodes), rn)['rel'] <= 1e-5
    assert np.abs(np.asarray(vels) - rv).max() <= 1e-4
    rtl, rrl = ref.vo_loss()
    assert np.allclose(tl.cpu().numpy(), rtl, rtol=2e-3, atol=1e-8) and np.allclose(rl.cpu().numpy(), rrl, rtol=2e-3, atol=1e-9)
    (tl2, rl2, nodes2, vels2, _), st2 = _general(g)
    assert (st2.steps_done, st2.tries_total, st2.reject_count) == (st.steps_done, st.tries_total, st.reject_count)
    assert np.abs(np.asarray(nodes) - np.asarray(nodes2)).max() <= 2e-6 and np.abs(np.asarray(vels) - np.asarray(vels2)).max() <= 2e-6


def test_rejected_tries_and_fixed_steps():
    g = synth.window(N=12)
    d = np.random.default_rng(1).standard_normal((g.N, 6)) * np.array([2, 2, 2, 0.8, 0.8, 0.8])
    g.init_nodes = lie.se3_retract(g.init_nodes.astype(np.float64), d).astype(np.float32)
    ref = po.SparseLM(g, np.float64, radius=1e6)
    for _ in range(5):
        ref.step()
    tl, rl, nodes, vels, _ = run_pvgo(*_args(g), radius=1e6, loss_weight=g.loss_weight, use_scheduler=False, max_steps=5)
    st = run_pvgo.last_state
    rejects = [h['rejects'] for h in ref.history]
    assert st.steps_done == 5 and st.tries_total == 5 + sum(rejects), (st.as_dict(), rejects)
    assert abs(st.loss - ref.history[-1]['loss']) <= 1e-3 * max(1.0, abs(ref.history[-1]['loss']))
    rn, _ = ref.aligned(g.init_nodes[0])
    assert po.rel_pose_error(np.asarray(nodes), rn)['rel'] <= 1e-4


def test_largest_window_with_loop_closure_and_band():
    g = synth.config2(N=16, band=3)
    from islam_b200.synth import _se3_mul, _se3_inv
    lc = np.array([[0, 15], [2, 11]])
    Zlc = _se3_mul(_se3_inv(g.gt_nodes[lc[:, 0]].astype(np.float64)), g.gt_nodes[lc[:, 1]].astype(np.float64)).astype(np.float32)
    g.links = np.concatenate([g.links, lc])
    g.vo_motions = np.concatenate([g.vo_motions, Zlc])
    ref = po.SparseLM(g, np.float64, solver='splu').run(steps=4)
    tl, rl, nodes, vels, _ = run_pvgo(*_args(g), radius=g.radius, loss_weight=g.loss_weight, use_scheduler=False, max_steps=4)
    st = run_pvgo.last_state
    assert st.steps_done == 4 and st.info == 0
    assert abs(st.loss - ref.history[-1]['loss']) <= 1e-4 * abs(ref.history[-1]['loss'])
    rn, _ = ref.aligned(g.init_nodes[0])
    assert po.rel_pose_error(np.asarray(nodes), rn)['rel'] <= 1e-5
    # smallest graph, and one beyond the fast path's size goes to the general path transparently
    g2 = synth.window(N=2)
    ref2 = po.SparseLM(g2, np.float64).run()
    _, _, n2, _, _ = run_pvgo(*_args(g2), radius=g2.radius, loss_weight=g2.loss_weight)
    assert po.rel_pose_error(np.asarray(n2), ref2.aligned(g2.init_nodes[0])[0])['rel'] <= 1e-5
    assert not ipvgo._small_ok(17, 16) and ipvgo._small_ok(16, 128) and not ipvgo._small_ok(16, 129)


def test_vo_loss_gradient_from_the_fused_launch():
    g = synth.window()
    vo = _t(g.vo_motions).cuda().requires_grad_(True)
    a = list(_args(g)); a[2] = vo
    tl, rl, nodes, vels, _ = run_pvgo(*a, radius=g.radius, loss_weight=g.loss_weight)
    loss_bp = torch.cat((rl, tl))                                # train.py:280-283
    assert loss_bp.requires_grad
    loss_bp.backward(torch.ones_like(loss_bp))
    ref = po.SparseLM(g, np.float64).run()
    gt, gr = po.vo_loss_grad(ref.nodes, ref.edges, ref.poses)
    got = vo.grad.cpu().numpy()
    assert np.abs(got[:, 6]).max() == 0
    assert np.abs(got[:, :6] - (gt + gr)).max() < 5e-3 * max(1e-3, np.abs(gt + gr).max())


def test_batch_of_windows_equals_window_by_window():
    B = 37
    gs = [synth.window(seed=s) for s in range(B)]
    stack = lambda k: _t(np.stack([getattr(g, k) for g in gs]))
    out = run_pvgo_batch(stack('init_nodes'), stack('init_vels'), stack('vo_motions'), _t(gs[0].links), stack('dts'),
                         stack('imu_drots'), stack('imu_dtrans'), stack('imu_dvels'), radius=gs[0].radius, loss_weight=gs[0].loss_weight)
    tl, rl, nodes, vels = out
    assert nodes.shape == (B, 9, 7) and vels.shape == (B, 9, 3) and tl.shape == (B, 8)
    states = run_pvgo_batch.last_states
    for b in (0, 5, 36):
        one = run_pvgo(*_args(gs[b]), radius=gs[b].radius, loss_weight=gs[b].loss_weight)
        st = run_pvgo.last_state
        assert (states[b].steps_done, states[b].tries_total) == (st.steps_done, st.tries_total)
        assert torch.equal(nodes[b], _t(np.asarray(one[2]))) and torch.equal(vels[b], one[3])          # same kernel: bitwise
        assert torch.equal(tl[b].cpu(), one[0].cpu())
        ref = po.SparseLM(gs[b], np.float64).run()
        assert po.rel_pose_error(nodes[b].numpy(), ref.aligned(gs[b].init_nodes[0])[0])['rel'] <= 1e-5


def _with_spec(g, slots, **kw):
    """run_pvgo with the speculative retry slots capped (0 = the plain sequential loop)."""
    os.environ['ISLAM_SMALL_SPEC'] = str(slots)
    try:
        out = run_pvgo(*_args(g), radius=kw.pop('radius', g.radius), loss_weight=g.loss_weight, **kw)
        return out, run_pvgo.last_state
    finally:
        del os.environ['ISLAM_SMALL_SPEC']


@pytest.mark.parametrize('case', ['plateau_storm', 'perturbed'])
def test_speculative_retries_are_bit_identical_to_the_sequential_loop(case):
    """csrc/small.cuh SM_SPEC: rejected tries are re-damped, factored and evaluated eight at a time, then consumed in order by
    the real controller.  Same decisions, same try counts and the SAME BITS as one try after the other."""
    if case == 'plateau_storm':
        g, kw = synth.window(), {}
    else:
        g = synth.window(N=12)
        d = np.random.default_rng(1).standard_normal((g.N, 6)) * np.array([2, 2, 2, 0.8, 0.8, 0.8])
        g.init_nodes = lie.se3_retract(g.init_nodes.astype(np.float64), d).astype(np.float32)
        kw = dict(radius=1e6, use_scheduler=False, max_steps=5)
    (tl0, rl0, n0, v0, _), s0 = _with_spec(g, 0, **dict(kw))
    for slots in (8, 3):
        (tl1, rl1, n1, v1, _), s1 = _with_spec(g, slots, **dict(kw))
        assert (s1.steps_done, s1.tries_total, s1.reject_count, s1.info) == (s0.steps_done, s0.tries_total, s0.reject_count, s0.info)
        assert s1.loss == s0.loss and s1.damping == s0.damping
        assert np.array_equal(np.asarray(n0), np.asarray(n1)) and np.array_equal(np.asarray(v0), np.asarray(v1))
        assert torch.equal(tl0, tl1) and torch.equal(rl0, rl1)
    if case == 'plateau_storm':
        assert s0.tries_total == s0.steps_done + 16        # the case does reject tries
