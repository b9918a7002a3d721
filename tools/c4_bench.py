"""BASELINE config 4 (50 000 poses, 2 000 loop closures => dense root, csrc/dense_root.cuh) per LM try, on 1 GPU or — under
torchrun — on N GPUs with the block-column-cyclic distributed root (islam_b200/dist.py).  CUDA events on the solver's
stream; the multi-GPU line is the max over ranks.

    python tools/c4_bench.py [--N 50000 --n-lc 2000 --tries 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/c4_bench.py
"""
import argparse, ctypes as C, os, sys, time
os.environ.setdefault('TORCH_NCCL_HIGH_PRIORITY', '1')      # see islam_b200/dist.py: look-ahead broadcasts
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from islam_b200 import synth, _lib

ap = argparse.ArgumentParser()
ap.add_argument('--N', type=int, default=50000)
ap.add_argument('--n-lc', type=int, default=2000)
ap.add_argument('--tries', type=int, default=3)
ap.add_argument('--one-gpu', action='store_true', help='all ranks on cuda:0 over gloo (functional check only)')
a = ap.parse_args()
world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
g = synth.config4(N=a.N, n_lc=a.n_lc)


def ev():
    return torch.cuda.Event(enable_timing=True)


if world == 1:
    from islam_b200.solver import PVGOSolver
    t0 = time.perf_counter()
    s = PVGOSolver(g.N, g.links)
    s.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight)
    s.set_state(g.init_nodes, g.init_vels)
    torch.cuda.synchronize()
    tc = time.perf_counter() - t0
    s.lm_reset(radius=g.radius, max_steps=a.tries + 1, use_scheduler=0)
    out = []
    for k in range(a.tries + 1):
        e0, e1 = ev(), ev()
        with torch.cuda.stream(s.stream):
            e0.record(); st = s.lm_step(); e1.record()
        torch.cuda.synchronize()
        out.append((e0.elapsed_time(e1), st.tries_total, st.loss, st.info))
    ms = [o[0] / max(1, o[1] - (out[i - 1][1] if i else 0)) for i, o in enumerate(out)]
    print(f'C4 N={g.N} closures={a.n_lc} 1 GPU: root {s.dims.root_pivots} variables, {s.dims.F} fronts / {s.dims.levels} levels, '
          f'setup {tc:.2f} s; ms per try {["%.1f" % m for m in ms]} (first = cold), losses {["%.6g" % o[2] for o in out]}, info {out[-1][3]}',
          flush=True)
else:
    import torch.distributed as dist
    from islam_b200.dist import ShardedPVGO
    lr = int(os.environ.get('LOCAL_RANK', '0'))
    dev = torch.device('cuda', 0 if a.one_gpu else lr)
    torch.cuda.set_device(dev)
    if a.one_gpu:
        dist.init_process_group('gloo')
    else:
        dist.init_process_group('nccl', device_id=dev)
    t0 = time.perf_counter()
    sh = ShardedPVGO(g.N, g.links, dev, exchange='nccl' if a.one_gpu else 'p2p')
    sh.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight)
    sh.set_state(g.init_nodes, g.init_vels)
    torch.cuda.synchronize()
    tc = time.perf_counter() - t0
    s = sh.s
    sh.lm_reset(radius=g.radius, max_steps=a.tries + 1, use_scheduler=0)
    st_ = C.c_void_p(s.stream.cuda_stream)
    rows = []
    for k in range(a.tries + 1):
        e = [ev() for _ in range(6)]
        dist.barrier()
        torch.cuda.synchronize()
        with torch.cuda.stream(s.stream):
            e[0].record()
            _lib.check(s.L.islam_pvgo_lm_try_begin(s._h, st_), 'begin')
            e[1].record()
            sh._allreduce(sh.shared)
            if sh.root_n:
                sh._allreduce(sh.root_R)
                sh._allreduce(sh.root_diag)
            e[2].record()
            _lib.check(s.L.islam_pvgo_lm_try_mid(s._h, st_), 'mid')
            e[3].record()
            if sh.root_n:
                sh._root_factor(st_)
            e[4].record()
            if sh.root_n:
                _lib.check(s.L.islam_pvgo_lm_try_mid2(s._h, st_), 'mid2')
            if sh.exchange == 'nccl':
                sh._allreduce(sh.sums)
            _lib.check(s.L.islam_pvgo_lm_try_end(s._h, st_), 'end')
            e[5].record()
        torch.cuda.synchronize()
        t = torch.tensor([e[i].elapsed_time(e[i + 1]) for i in range(5)] + [e[0].elapsed_time(e[5])], dtype=torch.float64,
                         device=dev if not a.one_gpu else 'cpu')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        state = s.lm_state()
        rows.append((t.tolist(), state.loss, state.info, state.tries_total))
    if rank == 0:
        print(f'C4 N={g.N} closures={a.n_lc} {world} GPUs{" (one device, gloo)" if a.one_gpu else ""}: root {s.dims.root_pivots} variables '
              f'(n = {sh.root_n}), {s.dims.n_shared_fronts} shared fronts, setup {tc:.2f} s', flush=True)
        for t, loss, info, tries in rows:
            print('   try: subtrees + partial root %.1f | all-reduce %.1f | shared fronts + diag %.1f | distributed root factor %.1f | '
                  'back-substitution + trial + control %.1f | total %.1f ms (max over ranks)   loss %.6g info %d tries %d'
                  % (*t, loss, info, tries), flush=True)
    dist.destroy_process_group()
