"""GPU parity tests (through the C ABI) of kernel families 1 and 2 and of the LM driver against the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from islam_b200 import synth
from islam_b200.solver import PVGOSolver
from oracle import pvgo_oracle as po

pytestmark = pytest.mark.gpu


def _solver(g, **kw):
    s = PVGOSolver(g.N, g.links, **kw)
    s.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight, reproj=_reproj_of(g, 'cuda'))
    s.set_state(g.init_nodes, g.init_vels)
    return s


def _dense_from_blocks(N, Hd, Ho, pairs):
    H = np.zeros((9 * N, 9 * N))
    for n in range(N):
        H[9 * n:9 * n + 9, 9 * n:9 * n + 9] = Hd[n]
    for p, (a, b) in enumerate(pairs):
        H[9 * a:9 * a + 9, 9 * b:9 * b + 9] = Ho[p]
        H[9 * b:9 * b + 9, 9 * a:9 * a + 9] = Ho[p].T
    return H


GRAPHS = {
    'C1': lambda: synth.config1(),
    'band8_300': lambda: synth.config2(N=300, band=8),
    'lc_400': lambda: synth.config4(N=400, n_lc=6, min_gap=50),
    'win9': lambda: synth.window(),
    'band3_57': lambda: synth.config2(N=57, band=3),
}


@pytest.mark.parametrize('name', list(GRAPHS))
def test_linearize_matches_oracle(name):
    g = GRAPHS[name]()
    s = _solver(g)
    s.linearize()
    res = [r.cpu().numpy() for r in s.residuals()]
    lm = po.SparseLM(g, np.float64)
    ref = lm._res()
    for a, b in zip(res, ref):
        assert np.abs(a - b).max() <= 2e-5 * max(1.0, np.abs(b).max()), name      # float32 evaluation of poses ~20 m
    Hd, Ho, gg, pairs = [t.cpu().numpy() for t in s.normal_equations()]
    H = _dense_from_blocks(g.N, Hd, Ho, pairs)
    Href, gref, _, _ = lm.assemble(ref)
    Href = Href.toarray()
    assert np.abs(H - H.T).max() == 0.0 or np.abs(H - H.T).max() < 1e-9 * np.abs(H).max()
    assert np.abs(H - Href).max() <= 5e-5 * np.abs(Href).max(), (name, np.abs(H - Href).max(), np.abs(Href).max())
    assert np.abs(gg - gref).max() <= 3e-4 * max(1.0, np.abs(gref).max())   # fp32 residuals (|t| ~ 20 m) through |J| ~ |t|


@pytest.mark.parametrize('name', list(GRAPHS))
def test_solve_matches_dense(name):
    """kernel family 2 alone: factor + solve the GPU's own H,g and compare with a dense float64 solve."""
    g = GRAPHS[name]()
    s = _solver(g)
    s.linearize()
    Hd, Ho, gg, pairs = [t.cpu().numpy() for t in s.normal_equations()]
    H = _dense_from_blocks(g.N, Hd, Ho, pairs)
    scale = 1.0 + 1e-4
    D, info = s.solve(scale)
    assert info == 0
    d = np.clip(np.diag(H), 1e-4, 1e32) * scale
    A = H.copy()
    A[np.arange(len(d)), np.arange(len(d))] = d
    Dref = np.linalg.solve(A, -gg.reshape(-1)).reshape(-1, 9)
    err = np.abs(D.cpu().numpy() - Dref).max() / np.abs(Dref).max()
    assert err < 1e-7, (name, err)


@pytest.mark.parametrize('opts', [dict(leaf_max=3, pivot_max=2), dict(leaf_max=16, pivot_max=4), dict(band_max=4),
                                  dict(leaf_max=1, pivot_max=1)])
@pytest.mark.parametrize('name', ['band8_300', 'lc_400', 'band3_57'])
def test_solve_matches_dense_with_other_front_shapes(name, opts):
    """The same check on plans with other front shapes: narrow / chained pivot sets, wide leaves, a band threshold that turns
    most of the band-8 edges into loop closures (a big ordinary root front with many children, update matrices in global
    memory), single-variable-triplet fronts."""
    g = GRAPHS[name]()
    s = _solver(g, **opts)
    s.linearize()
    Hd, Ho, gg, pairs = [t.cpu().numpy() for t in s.normal_equations()]
    H = _dense_from_blocks(g.N, Hd, Ho, pairs)
    scale = 1.0 + 1e-4
    D, info = s.solve(scale)
    assert info == 0
    d = np.clip(np.diag(H), 1e-4, 1e32) * scale
    A = H.copy()
    A[np.arange(len(d)), np.arange(len(d))] = d
    Dref = np.linalg.solve(A, -gg.reshape(-1)).reshape(-1, 9)
    err = np.abs(D.cpu().numpy() - Dref).max() / np.abs(Dref).max()
    assert err < 1e-7, (name, opts, err)


@pytest.mark.parametrize('name', ['C1', 'band8_300', 'lc_400'])
def test_lm_steps_match_oracle(name):
    """optimizer.step by optimizer.step: loss, damping, reject counts and the final (aligned) poses."""
    g = GRAPHS[name]()
    steps = 5
    ref = po.SparseLM(g, np.float64)
    s = _solver(g)
    s.lm_reset(radius=g.radius, max_steps=steps, use_scheduler=0)
    for k in range(steps):
        ref.step()
        st = s.lm_step()
        h = ref.history[-1]
        assert st.steps_done == k + 1
        assert st.reject_count == h['rejects'], (name, k, st.as_dict(), h)
        assert abs(st.loss - h['loss']) <= 1e-4 * max(1.0, abs(h['loss'])), (name, k, st.loss, h['loss'])
        assert abs(st.damping - h['damping']) <= 1e-12 * h['damping'], (name, k)
    n, v = s.align(g.init_nodes[0])
    rn, rv = ref.aligned(g.init_nodes[0])
    err = po.rel_pose_error(n.cpu().numpy(), rn)
    assert err['rel'] <= 1e-5, (name, err)                      # north_star tolerance: <= 1e-5 relative pose error
    assert np.abs(v.cpu().numpy() - rv).max() <= 1e-4 * max(1.0, np.abs(rv).max())


def test_window_with_scheduler_matches_oracle():
    """The shipped loop's size (batch_size=8 => 9 poses, run_kitti.sh:8) with StopOnPlateau(10, 3, 1e-3): the
    unweighted accept test makes step 2 burn its 16 rejects, which stops the scheduler."""
    g = synth.window()
    ref = po.SparseLM(g, np.float64).run()
    s = _solver(g)
    s.lm_reset(radius=g.radius, max_steps=10, patience=3, decreasing=1e-3, use_scheduler=1)
    st = s.lm_run()
    assert st.steps_done == len(ref.history)
    assert st.reject_count == ref.history[-1]['rejects']
    assert abs(st.loss - ref.history[-1]['loss']) <= 1e-4 * abs(ref.history[-1]['loss'])
    n, v = s.align(g.init_nodes[0])
    rn, rv = ref.aligned(g.init_nodes[0])
    assert po.rel_pose_error(n.cpu().numpy(), rn)['rel'] <= 1e-5


_MG_GRAPHS = {'band8': (lambda: synth.config2(N=600, band=8), 5),
              # 70 loop closures => a dense root of ~140 poses (~850 unknowns, 7 tile columns), csrc/dense_root.cuh
              'lcdense': (lambda: synth.config4(N=2500, n_lc=70, min_gap=100), 4),
              # the optional 5th residual group (pvgo.py:53-61): every rank evaluates the reprojection factors of the pairs it owns
              'band8rp': (lambda: _with_reproj(synth.config2(N=400, band=8)), 4)}


def _with_reproj(g, n_points=12, weight=2.0):
    g.extra['reproj'] = synth.reproj_data(g, n_points, weight=weight)
    g.loss_weight = tuple(g.loss_weight[:4]) + (weight,)
    return g


def _reproj_of(g, dev):
    """An object shaped like dense_ba.SparseReprojectionLoss (attributes only), or None."""
    rp = g.extra.get('reproj') if hasattr(g, 'extra') else None
    if rp is None:
        return None
    import types
    fx, fy, cx, cy = [float(v) for v in rp['K']]
    t = torch.as_tensor
    return types.SimpleNamespace(N=int(rp['point3d'].shape[1]), point3d=t(rp['point3d']).to(dev), target=t(rp['target']).to(dev),
                                 K=torch.tensor([fx, 0, cx, 0, fy, cy, 0, 0, 1], dtype=torch.float32).view(3, 3).to(dev),
                                 rgb2imu_pose=t(rp['rgb2imu']).to(dev))


def _mg_worker(rank, world, port, out, exchange, name='band8', one_gpu=False, lm_kw=None):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dev = torch.device('cuda', 0 if one_gpu else rank)
    torch.cuda.set_device(dev)
    if one_gpu:                                  # all ranks on ONE device: NCCL refuses that, gloo stages through the host
        dist.init_process_group('gloo', rank=rank, world_size=world)
    else:
        dist.init_process_group('nccl', device_id=dev)
    from islam_b200.dist import ShardedPVGO
    mk, steps = _MG_GRAPHS[name]
    g = mk()
    sh = ShardedPVGO(g.N, g.links, dev, exchange=exchange)
    sh.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight, reproj=_reproj_of(g, dev))
    sh.set_state(g.init_nodes, g.init_vels)
    sh.lm_reset(radius=g.radius, max_steps=steps, use_scheduler=0, **(lm_kw or {}))
    st = sh.lm_run()
    n, v = sh.get_state()
    if lm_kw:                                    # per-rank view of the LM state (failure tests)
        torch.save(dict(info=st.info, steps=st.steps_done, tries=st.tries_total, loss=st.loss), f'{out}.{rank}')
    if rank == 0:
        torch.save(dict(nodes=n.cpu(), vels=v.cpu(), loss=st.loss, steps=st.steps_done, rejects=st.reject_count, info=st.info,
                        root_n=sh.root_n, n_shared=sh.s.dims.n_shared_fronts), out)
    dist.destroy_process_group()


def _check_sharded(out, name, loss_rtol=1e-9):
    r = torch.load(out)
    mk, steps = _MG_GRAPHS[name]
    g = mk()
    s = _solver(g)
    s.lm_reset(radius=g.radius, max_steps=steps, use_scheduler=0)
    st = s.lm_run()
    n1, v1 = s.get_state()
    assert r['info'] == 0 and r['steps'] == steps and abs(r['loss'] - st.loss) <= loss_rtol * st.loss
    assert (r['nodes'] - n1.cpu()).abs().max().item() <= 1e-6
    ref = po.SparseLM(g, np.float64, solver='splu' if name == 'lcdense' else 'auto').run(steps=steps)
    assert po.rel_pose_error(r['nodes'].numpy(), ref.nodes)['rel'] <= 1e-5
    return r


@pytest.mark.parametrize('world,exchange', [(2, 'p2p'), (2, 'nccl'), (4, 'p2p')])
def test_sharded_lm_matches_single_gpu_and_oracle(world, exchange, tmp_path):
    """contiguous pose windows + one all-reduce of separator panels per try; identical decisions on all ranks.
    exchange='p2p': the trial sums travel through NVLink peer mailboxes inside the kernel that closes the try (no second
    collective); 'nccl': a 16-byte all-reduce instead."""
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    import torch.multiprocessing as mp
    out = str(tmp_path / 'mg.pt')
    mp.spawn(_mg_worker, args=(world, 29533 + world + (7 if exchange == 'nccl' else 0), out, exchange), nprocs=world, join=True)
    _check_sharded(out, 'band8')


@pytest.mark.parametrize('world,name', [(2, 'band8'), (2, 'lcdense'), (4, 'lcdense'), (2, 'band8rp')])
def test_sharded_ranks_on_one_gpu(world, name, tmp_path):
    """The N>1 path on a ONE-GPU box: `world` processes share cuda:0 and exchange through gloo (host-staged), so the
    sharded kernels — stage 1 / 2 of the separator fronts, the partial root, the block-column-cyclic dense-root Cholesky with
    its panel broadcasts (csrc/dense_root.cuh, include/islam_pvgo.h) — run wherever the GPU tests run."""
    import torch.multiprocessing as mp
    out = str(tmp_path / 'mg1.pt')
    port = 29571 + 10 * list(_MG_GRAPHS).index(name) + world          # one rendezvous port per case
    mp.spawn(_mg_worker, args=(world, port, out, 'nccl', name, True), nprocs=world, join=True)
    r = _check_sharded(out, name, loss_rtol=1e-8)
    assert (r['root_n'] > 0) == (name == 'lcdense')


@pytest.mark.parametrize('name', ['band8', 'lcdense'])
def test_sharded_failed_cholesky_is_seen_by_every_rank(name, tmp_path):
    """A factorisation that fails (every pivot diagonal clamped to -1: not SPD) must abandon the step on EVERY rank — with
    the distributed dense root only the owner of a block sees its non-positive pivot; the flag travels as a NaN in the
    exchanged trial sum — or the ranks would disagree on `continual` and the next collective would never complete."""
    import torch.multiprocessing as mp
    out = str(tmp_path / 'mgf.pt')
    mp.spawn(_mg_worker, args=(2, 29641 + len(name), out, 'nccl', name, True, dict(lm_max=-1.0)), nprocs=2, join=True)
    steps = _MG_GRAPHS[name][1]
    views = [torch.load(f'{out}.{r}') for r in range(2)]
    assert views[0] == views[1]
    assert views[0]['info'] == 1 and views[0]['steps'] == steps and views[0]['tries'] == steps
    g = _MG_GRAPHS[name][0]()
    assert np.array_equal(torch.load(out)['nodes'].numpy(), g.init_nodes)          # parameters untouched


@pytest.mark.parametrize('world', [2, 4])
def test_sharded_dense_root_over_nccl(world, tmp_path):
    """Config-4 structure on `world` GPUs: all-reduce of the partial root, then the distributed dense-root factorisation
    (NCCL broadcast of every factored block column from its owner)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    import torch.multiprocessing as mp
    out = str(tmp_path / 'mgd.pt')
    mp.spawn(_mg_worker, args=(world, 29611 + world, out, 'p2p', 'lcdense'), nprocs=world, join=True)
    r = _check_sharded(out, 'lcdense', loss_rtol=1e-8)
    assert r['root_n'] > 0


def test_config3_kitti_length_chain_with_gpu_preintegration():
    """BASELINE config 3: 4 541 poses, chain links, IMU deltas pre-integrated ON THE GPU from 100 Hz raw samples, then PVGO."""
    from islam_b200.imu_integrator import IMUModule
    N = 4541
    g = synth.config3(N=N)
    imu = synth.raw_imu(N)
    m = IMUModule(imu['accels'], imu['gyros'], imu['dts'], init=imu['init'], gravity=imu['gravity'],
                  rgb2imu_sync=imu['rgb2imu_sync'], device='cuda:0', denoise_accel=False, denoise_gyro=False)
    pos, rot, _, vel = m.integrate(0, N - 1, imu['init'], motion_mode=False)       # train.py:236  initial guess
    dtrans, drots, _, dvels = m.integrate(0, N - 1, imu['init'], motion_mode=True)  # train.py:244  factors
    g.init_nodes = np.concatenate([pos.numpy(), torch.as_tensor(rot).numpy()], 1).astype(np.float32)
    g.init_vels = vel.numpy().astype(np.float32)
    g.imu_drots, g.imu_dtrans, g.imu_dvels = torch.as_tensor(drots).numpy(), dtrans.numpy(), dvels.numpy()
    s = _solver(g)
    assert s.dims.band == 1 and s.dims.levels >= 9
    s.lm_reset(radius=g.radius, max_steps=10, use_scheduler=0)
    st = s.lm_run()
    ref = po.SparseLM(g, np.float64).run(steps=10)
    assert st.steps_done == 10 and st.info == 0
    for h, k in zip(ref.history[-1:], [st]):
        assert abs(k.loss - h['loss']) <= 1e-4 * abs(h['loss'])
    n, v = s.align(g.init_nodes[0])
    rn, rv = ref.aligned(g.init_nodes[0])
    assert po.rel_pose_error(n.cpu().numpy(), rn)['rel'] <= 1e-5


def test_config4_style_loop_closures_small_root():
    """BASELINE config 4's structure with a root small enough to be an ordinary front (chain + random loop closures)."""
    g = synth.config4(N=3000, n_lc=12, min_gap=100)
    s = _solver(g)
    assert 12 <= s.dims.root_pivots < 33
    s.lm_reset(radius=g.radius, max_steps=6, use_scheduler=0)
    st = s.lm_run()
    ref = po.SparseLM(g, np.float64, solver='splu').run(steps=6)
    assert st.steps_done == 6
    assert [h['rejects'] for h in ref.history][-1] == st.reject_count
    n, _ = s.align(g.init_nodes[0])
    rn, _ = ref.aligned(g.init_nodes[0])
    assert po.rel_pose_error(n.cpu().numpy(), rn)['rel'] <= 1e-5


def test_config4_dense_root_matches_oracle():
    """Many loop closures: their endpoints form a DENSE root (csrc/dense_root.cuh: tiled fp64 Cholesky of the Schur
    complement).  Kernel family 2 alone against a dense solve, then LM steps against the oracle."""
    g = synth.config4(N=2500, n_lc=70, min_gap=100)
    s = _solver(g)
    assert s.dims.root_pivots >= 100
    s.linearize()
    Hd, Ho, gg, pairs = [t.cpu().numpy() for t in s.normal_equations()]
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    N = g.N
    rows, cols, vals = [], [], []
    ar = np.arange(9)
    for n in range(N):
        rows.append(np.repeat(9 * n + ar, 9)); cols.append(np.tile(9 * n + ar, 9)); vals.append(Hd[n].ravel())
    for p_, (a, b) in enumerate(pairs):
        rows.append(np.repeat(9 * a + ar, 9)); cols.append(np.tile(9 * b + ar, 9)); vals.append(Ho[p_].ravel())
        rows.append(np.repeat(9 * b + ar, 9)); cols.append(np.tile(9 * a + ar, 9)); vals.append(Ho[p_].T.ravel())
    H = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(9 * N, 9 * N)).tocsc()
    scale = 1.0 + 1e-4
    d = np.clip(H.diagonal(), 1e-4, 1e32) * scale
    A = (H + sp.diags(d - H.diagonal())).tocsc()
    Dref = spla.splu(A).solve(-gg.reshape(-1)).reshape(-1, 9)
    D, info = s.solve(scale)
    assert info == 0
    assert np.abs(D.cpu().numpy() - Dref).max() <= 1e-7 * np.abs(Dref).max()
    steps = 4
    s.lm_reset(radius=g.radius, max_steps=steps, use_scheduler=0)
    ref = po.SparseLM(g, np.float64, solver='splu')
    for k in range(steps):
        ref.step()
        st = s.lm_step()
        assert st.reject_count == ref.history[-1]['rejects']
        assert abs(st.loss - ref.history[-1]['loss']) <= 1e-4 * abs(ref.history[-1]['loss'])
    n, _ = s.align(g.init_nodes[0])
    rn, _ = ref.aligned(g.init_nodes[0])
    assert po.rel_pose_error(n.cpu().numpy(), rn)['rel'] <= 1e-5


def _dense_reference_solve(s, g, scale):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    Hd, Ho, gg, pairs = [t.cpu().numpy() for t in s.normal_equations()]
    N = g.N
    rows, cols, vals = [], [], []
    ar = np.arange(9)
    for n in range(N):
        rows.append(np.repeat(9 * n + ar, 9)); cols.append(np.tile(9 * n + ar, 9)); vals.append(Hd[n].ravel())
    for p_, (a, b) in enumerate(pairs):
        rows.append(np.repeat(9 * a + ar, 9)); cols.append(np.tile(9 * b + ar, 9)); vals.append(Ho[p_].ravel())
        rows.append(np.repeat(9 * b + ar, 9)); cols.append(np.tile(9 * a + ar, 9)); vals.append(Ho[p_].T.ravel())
    H = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(9 * N, 9 * N)).tocsc()
    d = np.clip(H.diagonal(), 1e-4, 1e32) * scale
    A = (H + sp.diags(d - H.diagonal())).tocsc()
    return spla.splu(A).solve(-gg.reshape(-1)).reshape(-1, 9)


@pytest.mark.parametrize('N,n_lc,gap', [(600, 34, 40), (900, 52, 40), (1500, 95, 60), (3000, 210, 100)])
def test_dense_root_sizes_against_sparse_lu(N, n_lc, gap):
    """The dense-root kernels (csrc/dense_root.cuh: potrf with explicit inverse, DMMA panel solve, TMA-staged K = 128 update,
    multi-CTA back-substitution) on roots of different sizes — ragged last 64- and 128-column blocks, one to a dozen tile
    columns — against SciPy's sparse LU on the same damped normal equations."""
    g = synth.config4(N=N, n_lc=n_lc, min_gap=gap)
    s = _solver(g)
    assert s.dims.root_pivots >= 60
    s.linearize()
    scale = 1.0 + 1e-4
    Dref = _dense_reference_solve(s, g, scale)
    D, info = s.solve(scale)
    assert info == 0
    assert np.abs(D.cpu().numpy() - Dref).max() <= 1e-7 * np.abs(Dref).max()


def test_config4_large_dense_root_properties():
    """20 000 poses / 800 loop closures (root ~1 600 poses = 14 400 unknowns): no oracle at this size, so check the LM
    invariants — every accepted step lowers the (unweighted) loss, the linear solves succeed, the run is reproducible bit
    for bit."""
    g = synth.config4(N=20000, n_lc=800, min_gap=100)
    s = _solver(g)
    assert s.dims.root_pivots >= 1400
    out = []
    for rep in range(2):
        s.set_state(g.init_nodes, g.init_vels)
        s.lm_reset(radius=g.radius, max_steps=4, use_scheduler=0)
        losses = []
        for k in range(4):
            st = s.lm_step()
            assert st.info == 0
            losses.append(st.loss)
            assert st.loss <= st.last or st.reject_count >= 16
        out.append((losses, s.get_state()[0].cpu().numpy()))
    assert out[0][0][-1] < 0.2 * out[0][0][0] or out[0][0][-1] < out[0][0][0]
    # bitwise reproducible: the root's children extend-add colour by colour without atomics (csrc/pvgo.cu, dense_root.cuh)
    assert np.array_equal(out[0][1], out[1][1])
    assert out[0][0] == out[1][0]


def test_config4_full_size_against_oracle_fixture():
    """FULL BASELINE config 4 (50 000 poses, 51 999 edges, 2 000 loop closures => dense root of 24 519 unknowns) against the
    float64 CPU oracle: loss and reject count after each of 8 LM steps, poses after step 3 (far from convergence: the iterates
    themselves agree) and after step 8.  The oracle run (~80 s of CPU per step) is committed as a fixture,
    tests/golden/make_c4_golden.py; poses are compared on every 25th pose after align_to."""
    fx = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'c4_oracle_steps.npz'))
    steps, stride, mid = int(fx['steps']), int(fx['stride']), int(fx['mid'])
    g = synth.config4()
    assert g.N == int(fx['N']) and g.E == int(fx['E'])
    s = _solver(g)
    s.lm_reset(radius=g.radius, max_steps=steps, use_scheduler=0)
    # measured: losses agree to <= 1.2e-6 relative at every step; poses to 2.4e-5 relative (1.2 mm, 1.4e-5 rad) at steps 3 and 8.
    # The problem is nowhere near a fixed point there (it creeps for > 60 steps under heavy damping), so the float32 state /
    # linearisation (the reference's dtype) and the oracle's float64 iterates differ more than at convergence (C2: 6.7e-7).
    errs, GATE_MID, GATE_END = {}, 5e-5, 5e-5
    for k in range(steps):
        st = s.lm_step()
        assert st.info == 0 and st.reject_count == int(fx['rejects'][k])
        assert abs(st.loss - fx['losses'][k]) <= 1e-5 * abs(fx['losses'][k]), (k, st.loss, fx['losses'][k])
        if k + 1 in (mid, steps):
            n, v = s.align(g.init_nodes[0])
            rn, rv = (fx['nodes_mid'], fx['vels_mid']) if k + 1 == mid else (fx['nodes'], fx['vels'])
            err = po.rel_pose_error(n.cpu().numpy()[::stride], rn)
            errs[k + 1] = err
            assert np.abs(v.cpu().numpy()[::stride] - rv).max() <= 1e-3
    print('C4 full size, relative pose error vs the float64 oracle after step', errs)
    assert errs[mid]['rel'] <= GATE_MID and errs[steps]['rel'] <= GATE_END, errs


def test_rejected_tries_follow_the_oracle():
    """A badly perturbed initial guess makes Gauss-Newton overshoot: the second step burns 4 rejected tries before it is
    accepted.  The device-side roll-back, cumulative damping (A.diag += A.diag * damping per retry) and reject counting must
    follow SURVEY.md A.4 step by step."""
    from oracle import lie
    g = synth.config2(N=60, band=2)
    d = np.random.default_rng(1).standard_normal((g.N, 6)) * np.array([2, 2, 2, 0.8, 0.8, 0.8])
    g.init_nodes = lie.se3_retract(g.init_nodes.astype(np.float64), d).astype(np.float32)
    ref = po.SparseLM(g, np.float64, radius=1e6)
    s = _solver(g)
    s.lm_reset(radius=1e6, max_steps=5, use_scheduler=0)
    rejects = []
    for k in range(5):
        ref.step()
        st = s.lm_step()
        h = ref.history[-1]
        rejects.append(h['rejects'])
        assert st.reject_count == h['rejects'], (k, st.as_dict(), h)
        assert abs(st.loss - h['loss']) <= 1e-3 * max(1.0, abs(h['loss'])), (k, st.loss, h['loss'])
        assert abs(st.damping - h['damping']) <= 1e-9 * h['damping'], (k, st.damping, h['damping'])
    assert max(rejects) >= 2, rejects
    assert st.tries_total == 5 + sum(rejects)
    n, _ = s.align(g.init_nodes[0])
    rn, _ = ref.aligned(g.init_nodes[0])
    assert po.rel_pose_error(n.cpu().numpy(), rn)['rel'] <= 1e-4      # far from convergence, rotations ~1 rad: float32 state


def test_run_with_scheduler_stops_like_stop_on_plateau():
    g = synth.config2(N=200, band=4)
    ref = po.SparseLM(g, np.float64).run()                      # StopOnPlateau(10, 3, 1e-3)
    s = _solver(g)
    s.lm_reset(radius=g.radius, max_steps=10, patience=3, decreasing=1e-3, use_scheduler=1)
    st = s.lm_run()
    assert st.steps_done == len(ref.history) and st.continual == 0
    assert abs(st.loss - ref.history[-1]['loss']) <= 1e-4 * abs(ref.history[-1]['loss'])


def test_full_size_c2_properties():
    """BASELINE config 2 at full size (5 000 poses / 49 962 factors) through size-independent properties: parity with the
    oracle after the full 10 iterations, bitwise run-to-run determinism, and invariance to the order of the factor list."""
    g = synth.config2()
    s = _solver(g)

    def run(nodes):
        s.set_state(nodes, g.init_vels)
        s.lm_reset(radius=g.radius, max_steps=10, use_scheduler=0)
        st = s.lm_run()
        n, v = s.get_state()
        return st, n.cpu().numpy().astype(np.float64), v.cpu().numpy().astype(np.float64)

    st1, n1, v1 = run(g.init_nodes)
    st2, n2, v2 = run(g.init_nodes)
    assert st1.steps_done == 10 and st1.tries_total == st2.tries_total
    assert np.array_equal(n1, n2) and np.array_equal(v1, v2) and st1.loss == st2.loss          # deterministic gathers
    # the factor list is a set: a random permutation of the VO edges (links and measurements together) changes the
    # symbolic bookkeeping and every summation order, not the problem
    perm = np.random.default_rng(5).permutation(g.E)
    g2 = synth.config2()
    g2.links, g2.vo_motions = g.links[perm], g.vo_motions[perm]
    s2 = _solver(g2)
    s2.lm_reset(radius=g.radius, max_steps=10, use_scheduler=0)
    st3 = s2.lm_run()
    n3, v3 = [t.cpu().numpy().astype(np.float64) for t in s2.get_state()]
    # (float32 residuals summed in another order, ten iterations of a cond ~1e8 system: the loss agrees to ~1e-6 relative)
    assert st3.tries_total == st1.tries_total and abs(st3.loss - st1.loss) <= 3e-6 * st1.loss, (st3.loss, st1.loss)
    assert np.abs(n3 - n1).max() < 1e-4 and np.abs(v3 - v1).max() < 1e-4
    ref = po.SparseLM(g, np.float64).run(steps=10)
    s.set_state(g.init_nodes, g.init_vels)
    s.lm_reset(radius=g.radius, max_steps=10, use_scheduler=0)
    s.lm_run()
    n, _ = s.align(g.init_nodes[0])
    rn, _ = ref.aligned(g.init_nodes[0])
    err = po.rel_pose_error(n.cpu().numpy(), rn)
    assert err['rel'] <= 1e-5, err


def _check_steps(g, steps=3, tol=1e-5):
    ref = po.SparseLM(g, np.float64)
    s = _solver(g)
    s.linearize()
    Hd, Ho, gg, pairs = [t.cpu().numpy() for t in s.normal_equations()]
    Href, gref, _, _ = ref.assemble(ref._res())
    H = _dense_from_blocks(g.N, Hd, Ho, pairs)
    assert np.abs(H - Href.toarray()).max() <= 5e-5 * max(1e-6, np.abs(Href.toarray()).max())
    s.lm_reset(radius=g.radius, max_steps=steps, use_scheduler=0)
    for k in range(steps):
        ref.step()
        st = s.lm_step()
        assert st.reject_count == ref.history[-1]['rejects']
        assert abs(st.loss - ref.history[-1]['loss']) <= 1e-4 * max(1e-3, abs(ref.history[-1]['loss']))
    n, _ = s.align(g.init_nodes[0])
    rn, _ = ref.aligned(g.init_nodes[0])
    assert po.rel_pose_error(n.cpu().numpy(), rn)['rel'] <= tol


def test_edge_cases_minimal_empty_duplicate_and_reversed_edges():
    """Smallest graph (2 poses), no VO edges at all (IMU factors only), duplicated and reversed VO edges (the Hessian block
    of a pose pair then sums several factors; `links[:,0] > links[:,1]` flips the sign convention of pvgo.py:36-38)."""
    _check_steps(synth.window(N=2), steps=2)
    g = synth.config2(N=12, band=1)
    g.links = np.zeros((0, 2), np.int64)
    g.vo_motions = np.zeros((0, 7), np.float32)
    from oracle import lie
    d = np.random.default_rng(2).standard_normal((g.N, 6)) * 0.02          # the dead-reckoned guess satisfies the IMU factors
    g.init_nodes = lie.se3_retract(g.init_nodes.astype(np.float64), d).astype(np.float32)      # exactly: perturb it
    g.init_vels = (g.init_vels + np.random.default_rng(4).standard_normal((g.N, 3)).astype(np.float32) * 0.02)
    _check_steps(g, steps=2)
    g = synth.config2(N=40, band=3)
    rng = np.random.default_rng(3)
    dup = rng.choice(g.E, 25, replace=False)
    g.links = np.concatenate([g.links, g.links[dup]])
    g.vo_motions = np.concatenate([g.vo_motions, g.vo_motions[dup]])
    rev = rng.choice(g.E, 30, replace=False)                      # reverse: (j, i) with the inverse measurement
    g.links[rev] = g.links[rev][:, ::-1]
    g.vo_motions[rev] = lie.se3_inv(g.vo_motions[rev].astype(np.float64)).astype(np.float32)
    _check_steps(g, steps=3)
