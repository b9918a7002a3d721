"""Drop-in mirror of /root/reference/pvgo.py on top of the B200 kernels.

Same names, argument meaning, return order and error behaviour as the reference module:

    PoseVelGraph(nodes, vels)                      pvgo.py:15-23
        .forward(edges, poses, imu_drots, imu_dtrans, imu_dvels, dts)      pvgo.py:26-64
        .vo_loss(edges, poses) / .imu_loss(imu_drots, imu_dvels)           pvgo.py:67-78 / 95-111
        .align_to(target, idx=0)                                           pvgo.py:114-119
    run_pvgo(init_nodes, init_vels, vo_motions, links, dts, imu_drots, imu_dtrans, imu_dvels,
             device='cuda:0', radius=1e4, loss_weight=(1,1,1,1), reproj=None, target='vo')   pvgo.py:122-205

The reference builds a dense PyPose LM problem per call; here the graph structure (links) is analysed once and
cached, the information matrices collapse to the four scalars they are made of (pvgo.py:125-129), and the whole
`while scheduler.continual()` loop (pvgo.py:177-180) runs on the device.  There is no CPU path.
"""
import numpy as np
import torch

from ._lib import IslamError
from .solver import PVGOSolver

_SOLVER_POOL = {}          # structural key -> idle PVGOSolver handles (symbolic plan + device workspace)
_POOL_MAX = 8


def _plain(t):
    """LieTensor / Tensor / array -> plain torch.Tensor (keeps autograd history)."""
    if isinstance(t, torch.Tensor):
        return t.as_subclass(torch.Tensor) if type(t) is not torch.Tensor else t
    return torch.as_tensor(np.asarray(t))


def _wrap_like(ref, data, ltype_name):
    """Re-wrap as a LieTensor when the compat shim is loaded and the caller handed us LieTensors."""
    try:
        from . import pypose_compat as pp
    except Exception:      # pragma: no cover
        return data
    if isinstance(ref, pp.LieTensor):
        return pp.LieTensor(data, ltype=getattr(pp, ltype_name))
    return data


def _links_np(links):
    return np.ascontiguousarray(_plain(links).detach().cpu().numpy(), dtype=np.int64).reshape(-1, 2)


def acquire_solver(N, links, device):
    """Check a solver for this graph structure OUT of the pool (the symbolic analysis is what is expensive and is reused);
    a handle is owned by exactly one PoseVelGraph at a time, so two live graphs never share device state.
    release_solver() hands it back."""
    links_np = _links_np(links)
    dev = torch.device(device)
    if dev.type == 'cuda' and dev.index is None:
        dev = torch.device('cuda', torch.cuda.current_device())
    # cheap structural key (two vectorised checksums), confirmed by an exact comparison with the cached edge list:
    # a cryptographic hash of the edge list cost ~0.6 ms per run_pvgo call at 40 000 edges
    flat = links_np.reshape(-1)
    key = (int(N), str(dev), links_np.shape[0], int(flat.sum()), int(flat[::3].sum()), int(flat[1::5].sum()))
    idle = _SOLVER_POOL.get(key, [])
    for k, s in enumerate(idle):
        if np.array_equal(s.links, links_np):
            idle.pop(k)
            s._pool_key = key
            return s
    s = PVGOSolver(N, links_np, device=dev)
    s._pool_key = key
    return s


def release_solver(s):
    key = getattr(s, '_pool_key', None)
    if key is None:
        return
    if sum(len(v) for v in _SOLVER_POOL.values()) >= _POOL_MAX:       # evict the oldest idle handle
        for k in list(_SOLVER_POOL):
            if _SOLVER_POOL[k]:
                _SOLVER_POOL[k].pop(0)
                break
            del _SOLVER_POOL[k]
    _SOLVER_POOL.setdefault(key, []).append(s)


class _VoLoss(torch.autograd.Function):
    """vo_loss (pvgo.py:67-78): e = Log(P^-1 n1^-1 n2) with the nodes detached; autograd only reaches P.
    The gradient is returned the way PyPose's LieTensor backward does: left-tangent (6) padded to the 7-slot
    embedding, so upstream LieTensor ops (train.py:215) keep working."""

    @staticmethod
    def forward(ctx, P, solver):
        tl, rl, gt, gr = solver.vo_loss(P, with_grad=True)
        ctx.save_for_backward(gt, gr)
        ctx.dev = P.device
        return tl, rl

    @staticmethod
    def backward(ctx, g_tl, g_rl):
        gt, gr = ctx.saved_tensors
        g = g_tl.to(gt.device).unsqueeze(-1) * gt + g_rl.to(gr.device).unsqueeze(-1) * gr
        g = torch.cat([g, torch.zeros_like(g[:, :1])], dim=1)
        return g.to(ctx.dev), None


class _ImuLoss(torch.autograd.Function):
    """imu_loss (pvgo.py:95-111) evaluated on the caller's (possibly grad-carrying) imu_drots / imu_dvels
    (pvgo.py:149-150,188-189).  Gradients: d trans_loss / d dv = 2 adjvelerr; d rot_loss / d dR in PyPose's convention
    (left tangent, padded to the 4-slot SO3 embedding)."""

    @staticmethod
    def forward(ctx, drots, dvels, solver):
        tl, rl, gr, gv = solver.imu_loss(drots, dvels, with_grad=True)
        ctx.save_for_backward(gr, gv)
        ctx.devs = (drots.device, dvels.device)
        return tl, rl

    @staticmethod
    def backward(ctx, g_tl, g_rl):
        gr, gv = ctx.saved_tensors
        g_r = g_rl.to(gr.device).unsqueeze(-1) * gr
        g_r = torch.cat([g_r, torch.zeros_like(g_r[:, :1])], dim=1)
        g_v = g_tl.to(gv.device).unsqueeze(-1) * gv
        return g_r.to(ctx.devs[0]), g_v.to(ctx.devs[1]), None


class PoseVelGraph(torch.nn.Module):
    """pvgo.py:15-119.  Parameters live in the solver handle on the GPU (float32, as the reference).  The handle is checked
    out of a pool for the lifetime of the graph, so every live graph has its own device state."""

    def __init__(self, nodes, vels, reproj=None, links=None, device='cuda:0'):
        super().__init__()
        nodes_t, vels_t = _plain(nodes).detach(), _plain(vels).detach()
        assert nodes_t.size(0) == vels_t.size(0)                              # pvgo.py:19
        self._device = torch.device(device)
        self._init = (nodes_t, vels_t)
        self._nodes_ref = nodes
        self.reproj = reproj
        self.solver = None
        self._links_obj = None
        if links is not None:
            self._bind(links)

    def __del__(self):
        s, self.solver = getattr(self, 'solver', None), None
        if s is not None:
            try:
                release_solver(s)
            except Exception:      # interpreter shutdown
                pass

    def _bind(self, links):
        self.solver = acquire_solver(self._init[0].size(0), links, self._device)
        self._links_obj = links
        self.solver.set_state(*self._init)

    def _check_edges(self, edges):
        """The graph structure is fixed at construction; a different edge list is an error, not silently ignored."""
        if edges is None or edges is self._links_obj:
            return
        if not np.array_equal(_links_np(edges), self.solver.links):
            raise IslamError('this PoseVelGraph was built for a different edge list; construct a new graph')
        self._links_obj = edges

    def _ensure(self, edges, poses, imu_drots, imu_dtrans, imu_dvels, dts, loss_weight=None):
        """Stage the measurements of THIS call (the reference module is stateless w.r.t. its inputs, pvgo.py:26)."""
        if self.solver is None:
            self._bind(edges)
        else:
            self._check_edges(edges)
        if loss_weight is not None:
            self._loss_weight = tuple(float(x) for x in loss_weight)
        lw = getattr(self, '_loss_weight', (1.0, 1.0, 1.0, 1.0))
        self.solver.set_problem(_plain(poses), _plain(imu_drots), _plain(imu_dtrans), _plain(imu_dvels),
                                _plain(dts).reshape(-1), lw, reproj=self.reproj)

    @property
    def nodes(self):
        n, _ = self.solver.get_state()
        return _wrap_like(self._nodes_ref, n, 'SE3_type')

    @property
    def vels(self):
        return self.solver.get_state()[1]

    def forward(self, edges, poses, imu_drots, imu_dtrans, imu_dvels, dts):
        """Returns (pgerr, adjvelerr, imuroterr, transvelerr[, reprojerr]) — pvgo.py:61-64."""
        self._ensure(edges, poses, imu_drots, imu_dtrans, imu_dvels, dts)
        self.solver.linearize()
        return self.solver.residuals()

    def vo_loss(self, edges, poses):
        self._check_edges(edges)
        P = _plain(poses)
        if P.requires_grad:
            return _VoLoss.apply(P, self.solver)
        return self.solver.vo_loss(P)

    def imu_loss(self, imu_drots=None, imu_dvels=None):
        """pvgo.py:95-111 on the tensors passed in (None: the measurements staged by the last forward / run_pvgo)."""
        if imu_drots is None or imu_dvels is None:
            return self.solver.imu_loss()
        dr, dv = _plain(imu_drots), _plain(imu_dvels)
        if dr.requires_grad or dv.requires_grad:
            return _ImuLoss.apply(dr, dv, self.solver)
        return self.solver.imu_loss(dr, dv)

    def align_to(self, target, idx=0):
        if idx != 0:
            raise NotImplementedError('align_to is only used with idx=0 (pvgo.py:195)')
        n, v = self.solver.align(_plain(target).detach().reshape(7))
        return _wrap_like(self._nodes_ref, n, 'SE3_type'), v


def _small_ok(n_nodes, n_links):
    import os
    if os.environ.get('ISLAM_NO_SMALL') == '1':
        return False
    from . import _lib
    return bool(_lib.lib().islam_pvgo_small_supported(int(n_nodes), int(n_links)))


def _run_small(init_nodes, init_vels, vo_motions, links, dts, imu_drots, imu_dtrans, imu_dvels, device, radius, loss_weight,
               max_steps, patience, decreasing, use_scheduler, batch=None):
    from . import small
    N = init_nodes.shape[-2]
    B = 1 if batch is None else batch
    r = small.get_runner(N, links, device, B)
    P = _plain(vo_motions)
    need_grad = P.requires_grad and torch.is_grad_enabled()
    args = dict(nodes0=init_nodes, vels0=init_vels, Z=P, drot=imu_drots, dtrans=imu_dtrans, dvel=imu_dvels, dt=dts)
    states, nodes, vels, tl, rl, gt, gr = r.run(args, loss_weight, vo_P=None, with_grad=need_grad, radius=float(radius), lm_min=1e-4,
                                                max_steps=int(max_steps), patience=int(patience), decreasing=float(decreasing),
                                                use_scheduler=1 if use_scheduler else 0)
    for st in states:
        if st.info == 1:
            print('Linear solver failed. Breaking optimization step...')        # PyPose's message (A.4)
    if need_grad:
        Pd = P.to(device=r.device, dtype=torch.float32).reshape(B, -1, 7)         # differentiable: the gradient reaches the caller's tensor
        tl, rl = small.PrecomputedVoLoss.apply(Pd, tl, rl, gt, gr)
    if batch is None:
        nodes, vels, tl, rl = nodes[0], vels[0], tl[0], rl[0]
        run_pvgo.last_state = states[0]
    else:
        run_pvgo_batch.last_states = states
    return tl, rl, _wrap_like(init_nodes, nodes, 'SE3_type'), vels


def run_pvgo_batch(init_nodes, init_vels, vo_motions, links, dts, imu_drots, imu_dtrans, imu_dvels, device='cuda:0', radius=1e4,
                   loss_weight=(1, 1, 1, 1), max_steps=10, patience=3, decreasing=1e-3, use_scheduler=True):
    """B independent windows of identical structure in ONE launch, one CTA per window: init_nodes (B,N,7), init_vels (B,N,3),
    vo_motions (B,E,7), links (E,2) shared, dts (B,N-1), imu_* (B,N-1,.).  Returns (trans_loss (B,E), rot_loss (B,E),
    nodes (B,N,7) cpu, vels (B,N,3) cpu): what B calls of run_pvgo (pvgo.py:122-205, target='vo') return, stacked.
    (train.py optimises its windows one after the other only because each is initialised from the previous one; windows that
    are independent — several sequences, or a re-run over stored initial states — batch.)"""
    if not torch.cuda.is_available():
        raise IslamError('run_pvgo_batch needs a CUDA device: there is no CPU fallback')
    B, N = init_nodes.shape[0], init_nodes.shape[1]
    if not _small_ok(N, len(links)):
        raise IslamError('run_pvgo_batch covers windows of up to 16 poses / 128 edges; use run_pvgo for larger graphs')
    return _run_small(init_nodes, init_vels, vo_motions, links, dts, imu_drots, imu_dtrans, imu_dvels, device, radius, loss_weight,
                      max_steps, patience, decreasing, use_scheduler, batch=B)


def run_pvgo(init_nodes, init_vels, vo_motions, links, dts, imu_drots, imu_dtrans, imu_dvels,
             device='cuda:0', radius=1e4, loss_weight=(1, 1, 1, 1), reproj=None, target='vo',
             max_steps=10, patience=3, decreasing=1e-3, use_scheduler=True):
    """pvgo.py:122-205.  Returns (trans_loss, rot_loss, nodes [cpu, detached], vels [cpu, detached], covs)."""
    if not torch.cuda.is_available():
        raise IslamError('run_pvgo needs a CUDA device: there is no CPU fallback')
    n_links, n_nodes = len(links), len(init_nodes)
    # pvgo.py:125-131 — the information "matrices" are diagonal with these scalars
    vo_rot_infos = np.ones(n_links) * loss_weight[0] ** 2
    vo_trans_infos = np.ones(n_links) * loss_weight[0] ** 2
    imu_rot_infos = np.ones(n_nodes - 1) * loss_weight[2] ** 2
    imu_vel_infos = np.ones(n_nodes - 1) * loss_weight[1] ** 2
    transvel_infos = np.ones(n_nodes - 1) * loss_weight[3] ** 2
    if reproj is not None:
        reproj_infos = np.ones(n_nodes - 1) * (loss_weight[4] / reproj.N) ** 2        # pvgo.py:130-131

    if reproj is None and target == 'vo' and _small_ok(n_nodes, n_links):
        # the window sizes train.py uses (run_kitti.sh:8): the whole call is ONE kernel launch (csrc/small.cuh)
        out = _run_small(init_nodes, init_vels, vo_motions, links, dts, imu_drots, imu_dtrans, imu_dvels, device, radius,
                         loss_weight, max_steps, patience, decreasing, use_scheduler)
        covs = {'vo_rot': vo_rot_infos, 'imu_rot': imu_rot_infos, 'vo_trans': vo_trans_infos,
                'imu_vel': imu_vel_infos, 'transvel': transvel_infos}           # pvgo.py:199-203
        return out + (covs,)

    graph = PoseVelGraph(init_nodes, init_vels, reproj, links=links, device=device)
    # the whole call runs on the solver's stream (uploads, LM loop, outer loss, downloads): ordered ONCE after whatever the
    # caller's stream still has in flight (device-resident inputs), no cross-stream events per step; synchronised at the end
    graph.solver.stream.wait_stream(torch.cuda.current_stream(graph._device))
    with torch.cuda.stream(graph.solver.stream):
        # one host -> device transfer of the VO motions serves both the optimisation (detached, pvgo.py:146) and the outer
        # loss (autograd-connected, pvgo.py:186-187); .to() is differentiable, so the gradient still reaches the caller's tensor
        vo_dev = _plain(vo_motions).to(device=graph._device, dtype=torch.float32, non_blocking=True)
        graph._ensure(links, vo_dev.detach(), imu_drots, imu_dtrans, imu_dvels, dts, loss_weight)
        s = graph.solver
        # pvgo.py:169-180: LM(min=1e-4) + Cholesky + TrustRegion(radius) + StopOnPlateau(steps=10, patience=3, 1e-3)
        s.lm_reset(radius=float(radius), lm_min=1e-4, max_steps=int(max_steps), patience=int(patience),
                   decreasing=float(decreasing), use_scheduler=1 if use_scheduler else 0)
        st = s.lm_run()
        if st.info == 1:
            print('Linear solver failed. Breaking optimization step...')            # PyPose's message (A.4)
        elif st.info:
            raise IslamError(f'the device-side LM loop reported info={st.info} (see include/islam_pvgo.h)')

        if target == 'vo':                                                          # pvgo.py:186-189
            trans_loss, rot_loss = graph.vo_loss(links, vo_dev)
        elif target == 'imu':
            trans_loss, rot_loss = graph.imu_loss(imu_drots, imu_dvels)
        else:
            raise ValueError(f'unknown target {target!r}')

        nodes, vels = graph.align_to(_plain(init_nodes)[0])                         # pvgo.py:195
        # nodes.cpu(), vels.cpu() (pvgo.py:196-197) as two asynchronous copies into pinned memory and ONE synchronisation
        nodes_h = torch.empty(nodes.shape, dtype=nodes.dtype, pin_memory=True)
        vels_h = torch.empty(vels.shape, dtype=vels.dtype, pin_memory=True)
        nodes_h.copy_(_plain(nodes).detach(), non_blocking=True)
        vels_h.copy_(vels.detach(), non_blocking=True)
        torch.cuda.current_stream(graph._device).synchronize()
    nodes = _wrap_like(init_nodes, nodes_h, 'SE3_type')
    vels = vels_h
    covs = {'vo_rot': vo_rot_infos, 'imu_rot': imu_rot_infos, 'vo_trans': vo_trans_infos,
            'imu_vel': imu_vel_infos, 'transvel': transvel_infos}               # pvgo.py:199-203
    if reproj is not None:
        covs['reproj'] = reproj_infos
    run_pvgo.last_state = st
    return trans_loss, rot_loss, nodes, vels, covs
