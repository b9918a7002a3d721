#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its config[1]:
LM iterations/s (and factors/s) of the 5 000-pose / 49 962-factor synthetic PVGO (C2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over the graph: `run_pvgo`'s optimisation loop (/root/reference/pvgo.py:177-180)
run for exactly 10 `optimizer.step` calls from the dead-reckoned initial guess (scheduler bypassed so every step does
the same work).  `value` = LM iterations / second with inputs already resident in HBM; `e2e` = the same metric
through the reference-facing call `islam_b200.pvgo.run_pvgo(...)` fed HOST tensors (H2D of every input, D2H of
nodes / velocities / losses inside the timed region).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault('TORCH_NCCL_HIGH_PRIORITY', '1')      # islam_b200/dist.py: look-ahead broadcasts of the distributed dense root

LM_ITERS = 10                 # optimizer.step calls per step (StopOnPlateau's max steps, pvgo.py:172)
WORKLOAD = ('C2: 5000 poses / 39964 VO edges (band 8) + 4999 IMU pairs = 49962 factors, 284775 rows; '
            f'{LM_ITERS} fixed LM iterations per step, loss_weight (1,0.1,10,0.1), radius 1e4')
SURVEY_BYTES_PER_ITER = 62e6  # SURVEY.md 8d compact fp32 byte model of one LM iteration @C2
SURVEY_BYTES_FACTOR = 21.6e6  # read H (7.7 MB) + write L (13.9 MB): the factorisation's share of the above


def _peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.rows, self.p, self.index = [], None, index

    def __enter__(self):
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                       '--format=csv,noheader,nounits', '-lms', '100'],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None
        return self

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *a):
        if self.p:
            time.sleep(0.15)
            self.p.terminate()
            try:
                self.p.wait(timeout=2)
            except Exception:
                self.p.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


def _cpu_threads():
    """Threads the CPU port can actually use: its NumPy assembly is single-threaded, the LAPACK banded Cholesky runs on
    the BLAS thread pool."""
    try:
        from threadpoolctl import threadpool_info
        n = [int(i.get('num_threads', 1)) for i in threadpool_info() if i.get('user_api') == 'blas']
        return max(n) if n else 1
    except Exception:
        return 1


def _graph():
    from islam_b200 import synth
    return synth.config2()


# ---------------------------------------------------------------------------------------------------- reference arm
def time_oracle(g, iters, dtype=np.float64):
    """The CPU restatement (oracle.SparseLM: same normal equations, banded LAPACK Cholesky) for `iters` LM steps."""
    from oracle import pvgo_oracle as po
    lm = po.SparseLM(g, dtype)
    t0 = time.perf_counter()
    lm.run(steps=iters)
    return time.perf_counter() - t0, lm


def run_reference(args, rank, world):
    """The reference arm: the CPU port of the path (oracle.SparseLM; PyPose itself is absent) on the box's host
    cores, on the SAME config as our arm: every step is LM_ITERS optimizer.step calls on C2 from the same initial guess."""
    if rank != 0:
        return
    g = _graph()
    for _ in range(min(args.warmup, 1)):
        time_oracle(g, 1)
    t = 0.0
    for _ in range(args.steps):
        dt, _lm = time_oracle(g, LM_ITERS)
        t += dt
    its = LM_ITERS * args.steps / t
    cores = _cpu_threads()
    out = {
        'impl': 'reference', 'metric': 'LM iterations/s on the 5k-pose PVGO (C2)', 'value': its, 'unit': 'LM it/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t / args.steps,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'factors_per_s': its * g.factors, 'residual_rows_per_s': its * g.rows,
        'config': {'workload': WORKLOAD, 'parallelism': 'host CPU'},
        'cpu_baseline': {'value': its, 'unit': 'LM it/s', 'factors_per_s': its * g.factors, 'cores': cores, 'kind': 'port',
                         'sample': f'{LM_ITERS} LM iterations of C2 per step x {args.steps} steps (oracle.SparseLM float64: NumPy assembly + '
                                   f'LAPACK banded Cholesky; PyPose itself is absent and its dense algorithm needs 324 GB at C2)'},
        'e2e': {'value': its, 'unit': 'LM it/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    _emit(out)


def _pose_error_vs_fixture(nodes, vels=None):
    """Parity carried by the bench line: relative pose error (after align_to, pvgo.py:195) against the float64 CPU oracle's
    solution of the same graph after the same 10 iterations (tests/golden/c2_oracle_final.npz, made by make_c2_golden.py).
    `nodes` are already aligned.  Same definition as the parity tests: ||X - Xref||_F / ||Xref||_F on the 7-vector storage
    with the quaternion sign made canonical (w >= 0)."""
    p = os.path.join(ROOT, 'tests', 'golden', 'c2_oracle_final.npz')
    if not os.path.exists(p):
        return None
    fx = np.load(p)
    ref = fx['nodes'].astype(np.float64).copy()
    a = np.asarray(nodes, np.float64).copy()
    for x in (a, ref):
        x[:, 3:] = np.where(x[:, 6:7] < 0, -x[:, 3:], x[:, 3:])
    out = {'rel': float(np.linalg.norm(a - ref) / np.linalg.norm(ref)),
           'max_trans_m': float(np.max(np.linalg.norm(a[:, :3] - ref[:, :3], axis=1))),
           'gate': 1e-5, 'fixture': 'tests/golden/c2_oracle_final.npz (float64 CPU oracle, same 10 steps, aligned)'}
    if vels is not None:
        out['max_vel_mps'] = float(np.max(np.abs(np.asarray(vels, np.float64) - fx['vels'])))
    return out


def _align_np(nodes, vels, target):
    """align_to (pvgo.py:114-119) on host arrays, for the sharded runs whose gathered state comes back unaligned."""
    from islam_b200 import synth
    n, v, t = np.asarray(nodes, np.float64), np.asarray(vels, np.float64), np.asarray(target, np.float64)
    T = synth._se3_mul(t, synth._se3_inv(n[0]))
    q = synth._qmul(t[3:], synth._qinv(n[0, 3:]))
    return synth._se3_mul(T[None], n), synth._qrot(q[None], v)


# ---------------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device — the B200 kernels are the only implementation (no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from islam_b200.solver import PVGOSolver
    from islam_b200 import pvgo as ipvgo

    g = _graph()
    F = g.factors
    sharded = world > 1
    if sharded:
        # strong scaling: ONE C2 graph, contiguous pose windows, one all-reduce of separator panels per LM try
        from islam_b200.dist import ShardedPVGO
        sh = ShardedPVGO(g.N, g.links, dev)
        s = sh.s
    else:
        sh = None
        s = PVGOSolver(g.N, g.links, device=dev)
    s.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight)
    nodes0 = torch.as_tensor(g.init_nodes, device=dev)
    vels0 = torch.as_tensor(g.init_vels, device=dev)

    def one_step():
        s.set_state(nodes0, vels0)                    # D2D: inputs are resident in HBM
        s.lm_reset(radius=g.radius, max_steps=LM_ITERS, use_scheduler=0)
        return sh.lm_run() if sharded else s.lm_run()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        st = one_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tries = 0
    with ClockSampler(local_rank) as clk:
        with torch.cuda.stream(s.stream):
            e0.record()
            for _ in range(args.steps):
                st = one_step()
                tries += st.tries_total
            e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # parity of THIS run's result (the state the last timed step left behind), after align_to
    if sharded:
        n_fin, v_fin = sh.get_state()
        n_al, v_al = _align_np(n_fin.cpu().numpy(), v_fin.cpu().numpy(), g.init_nodes[0])
    else:
        n_al, v_al = s.align(nodes0[0])
        n_al, v_al = n_al.cpu().numpy(), v_al.cpu().numpy()
    pose_err = _pose_error_vs_fixture(n_al, v_al)
    iters_total = LM_ITERS * args.steps                  # one graph, however many GPUs share it
    value = iters_total / (ms * 1e-3)
    launches = tries * (7 + s.dims.levels + s.dims.bs_launches)   # begin_try, factors, assemble, begin_step, one factor launch per level, back-substitution (top levels chained in one launch), retract, trial factors, end_try

    # ---- per-phase device time of one try (CUDA events on the solver's stream) -> roofline of the dominant kernel
    one_step()
    s.set_state(nodes0, vels0)
    s.lm_reset(radius=g.radius, max_steps=LM_ITERS, use_scheduler=0)
    peak, peak_src = _peaks()
    per_iter_s = ms * 1e-3 / (LM_ITERS * args.steps)
    whole = {'achieved': SURVEY_BYTES_PER_ITER / per_iter_s / 1e9, 'frac': SURVEY_BYTES_PER_ITER / per_iter_s / 1e9 / peak}
    if sharded:
        # the phase split needs the single-GPU profiling hook; across ranks only the whole try is timed
        roofline = {'bound': 'hbm', 'kernel': 'whole LM try (linearise + factorisation + all-reduce + back-substitution + trial), max over ranks',
                    'achieved': whole['achieved'], 'peak': peak, 'unit': 'GB/s', 'frac': whole['frac'], 'traffic': None,
                    'peak_source': peak_src, 'algorithmic_bytes_per_iteration': SURVEY_BYTES_PER_ITER,
                    'avg_launch_us': per_iter_s * 1e6, 'whole_iteration': whole}
    else:
        ph = [s.profile_try() for _ in range(LM_ITERS)]
        fac_ms = float(np.mean([p['factor'] for p in ph]))
        per_launch_s = fac_ms * 1e-3 / s.dims.levels
        achieved = SURVEY_BYTES_FACTOR / (fac_ms * 1e-3) / 1e9
        roofline = {'bound': 'hbm', 'kernel': f'k_factor3 x{s.dims.levels} (multifrontal fp64 Cholesky of one LM try)',
                    'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None,
                    'peak_source': peak_src, 'algorithmic_bytes_per_factorisation': SURVEY_BYTES_FACTOR,
                    'avg_launch_us': per_launch_s * 1e6,
                    'phases_ms': {k: float(np.mean([p[k] for p in ph])) for k in ph[0]}, 'whole_iteration': whole}
    # `traffic` cannot be measured live (DRAM byte counters need ncu): it is the per-launch dram__bytes of the committed
    # `ncu --set full` capture of this same command (profiles/traffic.json names the capture); stated as such in the line
    tr = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tr) and not sharded:
        try:
            tj = json.load(open(tr))
            roofline['traffic'] = tj.get('k_factor_level_bytes_per_launch')
            roofline['traffic_source'] = 'ncu-derived, not live: ' + str(tj.get('source', 'profiles/traffic.json'))
        except Exception:
            pass

    # ---- e2e: the reference-facing call with HOST tensors (pinned), H2D + D2H inside the timed region
    host = {k: torch.as_tensor(getattr(g, k)).pin_memory() for k in
            ('init_nodes', 'init_vels', 'vo_motions', 'dts', 'imu_drots', 'imu_dtrans', 'imu_dvels')}
    links = torch.as_tensor(g.links)

    def e2e_step():
        if sharded:
            sh.set_problem(host['vo_motions'], host['imu_drots'], host['imu_dtrans'], host['imu_dvels'], host['dts'],
                           g.loss_weight)
            sh.set_state(host['init_nodes'], host['init_vels'])
            sh.lm_reset(radius=g.radius, max_steps=LM_ITERS, use_scheduler=0)
            st_ = sh.lm_run()
            n, v = sh.get_state()
            return st_.loss, n.cpu(), v.cpu()
        tl, rl, n, v, _ = ipvgo.run_pvgo(host['init_nodes'], host['init_vels'], host['vo_motions'], links, host['dts'],
                                         host['imu_drots'], host['imu_dtrans'], host['imu_dvels'], device=dev,
                                         radius=g.radius, loss_weight=g.loss_weight, use_scheduler=False,
                                         max_steps=LM_ITERS)
        return float(tl.sum().item() + rl.sum().item()), n, v

    for _ in range(2):
        e2e_step()
    barrier()
    n_e2e = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = LM_ITERS * n_e2e / float(te.item())
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = (g.N * 7 + g.N * 3 + 2 * g.E) * 4
    e2e = {'value': e2e_val, 'unit': 'LM it/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
           'api': ('islam_b200.pvgo.run_pvgo (mirror of reference pvgo.py:122-205), host tensors in, host tensors out' if not sharded
                   else 'islam_b200.dist.ShardedPVGO set_problem/set_state/lm_run/get_state, host tensors in and out')}

    out = {
        'metric': 'LM iterations/s on the 5k-pose PVGO (C2)', 'value': value, 'unit': 'LM it/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32 residuals/Jacobians, f64 normal equations + Cholesky',
        'data': 'synthetic', 'factors_per_s': value * F, 'residual_rows_per_s': value * g.rows,
        'config': {'workload': WORKLOAD,
                   'parallelism': 'single GPU' if world == 1 else
                   f'{world} contiguous pose windows of ONE C2 graph; per LM try one NCCL all-reduce of the shared separator panels '
                   f'({s.dims.n_shared_fronts} fronts, {s.dims.shared_doubles * 8 / 1e6:.2f} MB); the trial sums travel through NVLink peer mailboxes inside the kernel that closes the try',
                   'l2': (lambda d: f'no explicit flush: one LM iteration streams {(d.L_doubles * 8 + d.U_doubles * 10 + 648 * (d.N + d.P) + 336 * d.E) / 1e6:.0f} MB '
                                    f'(L {d.L_doubles * 8 / 1e6:.0f} MB + U {d.U_doubles * 8 / 1e6:.0f} MB + destination maps {d.U_doubles * 2 / 1e6:.0f} MB + '
                                    f'J^T W J blocks {648 * (d.N + d.P) / 1e6:.0f} MB + per-factor products {336 * d.E / 1e6:.0f} MB), more than the 126 MB L2')(s.dims)},
        'clocks': clk.summary(), 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline,
        'lm': {'final_loss': st.loss, 'steps_done': st.steps_done, 'tries': st.tries_total, 'info': st.info,
               'rel_pose_error_vs_oracle': pose_err},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        t_cpu, _ = time_oracle(g, 10)
        out['cpu_baseline'] = {'value': 10 / t_cpu, 'unit': 'LM it/s', 'factors_per_s': 10 / t_cpu * F, 'cores': _cpu_threads(), 'kind': 'port',
                               'sample': '10 LM iterations of C2 (oracle.SparseLM float64: single-threaded NumPy assembly + LAPACK banded '
                                         'Cholesky on the BLAS thread pool, same normal equations; literal dense PyPose needs 324 GB at C2)'}
    if not args.no_extras:
        try:
            from islam_b200.dist import ShardedPVGO as _Sh
            extras = _other_configs(rank, world, dev, _Sh)
        except Exception as exc:                                # never lose the headline to an extra
            extras = {'error': f'{type(exc).__name__}: {exc}'}
        out['other_configs'] = extras
    if rank == 0:
        _emit(out)
    if world > 1:
        sh.release_graph()
        torch.cuda.synchronize()
        dist.destroy_process_group()


def _other_configs(rank, world, dev, sh_cls):
    """The other BASELINE.json configs that have a story of their own, measured beside the headline (never part of `value`):
    C5's back-end share — run_pvgo on the shipped 9-pose window, host tensors in and out (N = 1 only) — and C4, 50 000 poses
    with 2 000 loop closures (dense root of 24 519 unknowns): ms per LM iteration, on N GPUs with the root factored by all
    ranks together (islam_b200/dist.py).  CUDA events / wall clock after warm-up, max over ranks."""
    import torch
    import torch.distributed as dist
    from islam_b200 import synth
    out = {}
    if world == 1:
        from islam_b200 import pvgo as ipvgo
        w = synth.window()
        t = lambda a_: torch.as_tensor(a_).pin_memory()
        a = [t(w.init_nodes), t(w.init_vels), t(w.vo_motions), torch.as_tensor(w.links), t(w.dts), t(w.imu_drots), t(w.imu_dtrans),
             t(w.imu_dvels)]
        for _ in range(10):
            ipvgo.run_pvgo(*a, device=dev, radius=w.radius, loss_weight=w.loss_weight)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(100):
            ipvgo.run_pvgo(*a, device=dev, radius=w.radius, loss_weight=w.loss_weight)
        torch.cuda.synchronize()
        st = ipvgo.run_pvgo.last_state
        out['C5_window_run_pvgo'] = {'ms_per_call': (time.perf_counter() - t0) / 100 * 1e3, 'poses': int(w.N), 'lm_steps': st.steps_done,
                                     'tries': st.tries_total, 'api': 'run_pvgo, host tensors in and out, StopOnPlateau (pvgo.py:122-205)'}
    g4 = synth.config4()
    tries = 3
    if world == 1:
        from islam_b200.solver import PVGOSolver
        s4 = PVGOSolver(g4.N, g4.links, device=dev)
        step = s4.lm_step
    else:
        s4w = sh_cls(g4.N, g4.links, dev)
        s4 = s4w.s

        def step():
            s4w.lm_try()
            return s4.lm_state()
    (s4w if world > 1 else s4).set_problem(g4.vo_motions, g4.imu_drots, g4.imu_dtrans, g4.imu_dvels, g4.dts, g4.loss_weight)
    (s4w if world > 1 else s4).set_state(g4.init_nodes, g4.init_vels)
    (s4w if world > 1 else s4).lm_reset(radius=g4.radius, max_steps=tries + 1, use_scheduler=0)
    st = step()                                           # warm-up (first touch of 4.8 GB of factor, NCCL channels)
    n0 = st.tries_total
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s4.stream):
        e0.record()
        for _ in range(tries):
            st = step()
        e1.record()
    torch.cuda.synchronize()
    tm = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_it = float(tm.item()) / max(1, st.tries_total - n0)
    n_root = int(s4.dims.max_cols)                        # the dense root is by far the widest front
    peak = 63 * 2 * 148 * 1.965e9 / 1e12                  # DMMA: 63 FMA/clk/SM measured (tools/pipe_bench.cu) x 148 SMs x 1.965 GHz
    tfl = (n_root ** 3 / 3) / (ms_it * 1e-3) / 1e12
    out['C4_dense_root_roofline'] = {'bound': 'tensor (fp64 DMMA)', 'unknowns': n_root, 'flops_per_factorisation': n_root ** 3 / 3,
                                     'achieved': tfl, 'peak': peak * world, 'unit': 'TFLOP/s', 'frac': tfl / (peak * world),
                                     'note': 'n^3/3 over the WHOLE LM iteration (subtrees, assembly, exchange, back-substitution '
                                             'included); peak = own DMMA micro-benchmark per GPU x n_gpus, MEASURED_PEAKS has no fp64 entry'}
    out['C4_loop_closures'] = {'ms_per_lm_iteration': ms_it, 'n_gpus': world,
                               'poses': int(g4.N), 'loop_closures': 2000, 'fronts': int(s4.dims.F),
                               'loss_after': st.loss, 'info': st.info,
                               'how': 'dense root factored block-column-cyclically by all ranks, NCCL broadcast of each factored block'
                                      if world > 1 else 'dense root on one GPU (csrc/dense_root.cuh)'}
    return out



def _emit(obj):
    """The ONE JSON line goes to the real stdout; everything libraries print meanwhile (NCCL banner ...) went to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + '\n').encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the C4 / C5 side measurements (other_configs)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
