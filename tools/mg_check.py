"""Multi-GPU check (launch with torchrun): sharded LM on C2 / a small band graph vs the single-GPU solver and the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from islam_b200 import synth
from islam_b200.dist import ShardedPVGO
from islam_b200.solver import PVGOSolver

rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
for name, g, steps, ex in (('band8_600', synth.config2(N=600, band=8), 5, 'p2p'), ('C2', synth.config2(), 10, 'nccl'), ('C2', synth.config2(), 10, 'p2p')):
    sh = ShardedPVGO(g.N, g.links, dev, exchange=ex)
    sh.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight)
    for rep in range(3):
        sh.set_state(g.init_nodes, g.init_vels)
        sh.lm_reset(radius=g.radius, max_steps=steps, use_scheduler=0)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        st = sh.lm_run()
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
    n, v = sh.get_state()
    if rank == 0:
        s1 = PVGOSolver(g.N, g.links, device=dev)
        s1.set_problem(g.vo_motions, g.imu_drots, g.imu_dtrans, g.imu_dvels, g.dts, g.loss_weight)
        s1.set_state(g.init_nodes, g.init_vels)
        s1.lm_reset(radius=g.radius, max_steps=steps, use_scheduler=0)
        st1 = s1.lm_run()
        n1, v1 = s1.get_state()
        d = (n - n1).abs().max().item()
        print(f'[{name}] exchange={ex} world={world} shared_fronts={sh.s.dims.n_shared_fronts} shared_MB={sh.s.dims.shared_doubles*8/1e6:.2f} '
              f'steps={st.steps_done} tries={st.tries_total} loss={st.loss:.9f} (1-GPU {st1.loss:.9f}) '
              f'max|nodes-nodes_1gpu|={d:.3e} ms/try={1e3*dt/max(1,st.tries_total):.3f}', flush=True)
        from oracle import pvgo_oracle as po
        ref = po.SparseLM(g, np.float64).run(steps=steps)
        print(f'[{name}] parity vs oracle (unaligned, same gauge path):', po.rel_pose_error(n.cpu().numpy(), ref.nodes), flush=True)
dist.destroy_process_group()
