"""Multi-GPU PVGO: contiguous pose windows, one process per GPU, torch.distributed (NCCL over NVLink) for the exchange.

SURVEY.md section 8e: the top log2(G) levels of the nested-dissection tree are the cuts between windows.  Every rank
owns the factors that touch its private variables, eliminates its own subtree, and contributes partial frontal matrices
of the shared (separator) fronts; ONE NCCL all-reduce per LM try sums those (plus the partial linearisation loss), after
which every rank factors the few shared fronts redundantly and back-substitutes its own window.  The two doubles every
rank needs for the common accept / roll-back decision (trial loss, quality term) do not go through a second collective:
the kernel that closes the try stores them straight into the peers' mailboxes over NVLink (CUDA IPC peer memory) and
sums the G messages in rank order (exchange='p2p', the default; exchange='nccl' keeps a 16-byte all-reduce instead).

A DENSE loop-closure root (BASELINE config 4) is the one part with enough arithmetic to shard: it is assembled by an
all-reduce of the ranks' shares and factored by all ranks together, 1-D block-column-cyclic, every factored 128-column
block broadcast from its owner with one block of look-ahead on a high-priority stream (_root_factor).
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import IslamError, LMState
from .solver import PVGOSolver


def _wrap(ptr, n, device):
    """A float64 tensor view over library-owned device memory (no copy), via the CUDA array interface."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {'shape': (int(n),), 'typestr': '<f8', 'data': (int(ptr), False), 'version': 3}
    return torch.as_tensor(h, device=device)


class ShardedPVGO:
    """One graph sharded over `world` ranks (world a power of two).  Collectives go through `group`; with the gloo
    backend (CPU test rigs) the buffers are staged through host memory."""

    def __init__(self, N, links, device, rank=None, world=None, group=None, exchange='p2p'):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        if self.world & (self.world - 1):
            raise IslamError('the number of windows must be a power of two')
        self.s = PVGOSolver(N, links, device=device, n_parts=self.world, part=self.rank)
        L, h = self.s.L, self.s._h
        p, n = C.c_void_p(), C.c_int64()
        _lib.check(L.islam_pvgo_shared_buffer(h, C.byref(p), C.byref(n)), 'islam_pvgo_shared_buffer')
        self.shared = _wrap(p.value, n.value, self.s.device)
        _lib.check(L.islam_pvgo_sums_buffer(h, C.byref(p), C.byref(n)), 'islam_pvgo_sums_buffer')
        self.sums = _wrap(p.value, n.value, self.s.device)
        # dense loop-closure root (BASELINE config 4): factored by all ranks together, see lm_try
        rp, rn, rld, dp, blk = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_void_p(), C.c_int32()
        _lib.check(L.islam_pvgo_root_buffers(h, C.byref(rp), C.byref(rn), C.byref(rld), C.byref(dp), C.byref(blk)),
                   'islam_pvgo_root_buffers')
        self.root_n, self.root_ld, self.root_block = rn.value, rld.value, blk.value
        if self.root_n:
            self.root_R = _wrap(rp.value, self.root_n * self.root_ld, self.s.device)
            self.root_diag = _wrap(dp.value, self.root_n, self.s.device)
            self._root_owner = [L.islam_pvgo_root_owner(h, k0) for k0 in range(0, self.root_n, self.root_block)]
            self._root_src = [o if group is None else dist.get_global_rank(group, o) for o in self._root_owner]
        parts = np.zeros(3 * N, np.int32)
        _lib.check(L.islam_pvgo_var_parts(h, parts.ctypes.data), 'islam_pvgo_var_parts')
        self.var_parts = parts = parts.reshape(N, 3)            # [tau, phi, v] of every pose
        # a pose / velocity is reported by the rank that solves it: the owner of a private variable, else rank 0
        pose_owner = np.where((parts[:, :2] >= 0).any(1), parts[:, :2].max(1), 0)
        vel_owner = np.where(parts[:, 2] >= 0, parts[:, 2], 0)
        self._mine_pose = torch.as_tensor(pose_owner == self.rank, device=self.s.device)
        self._mine_vel = torch.as_tensor(vel_owner == self.rank, device=self.s.device)
        self._nccl = dist.get_backend(group) == 'nccl'
        if exchange not in ('p2p', 'nccl'):
            raise IslamError("exchange must be 'p2p' or 'nccl'")
        self.exchange = exchange
        if exchange == 'p2p':
            # every rank exports the CUDA IPC handle of its mailbox; the table of all handles goes back into the library
            buf = (C.c_ubyte * 64)()
            _lib.check(L.islam_pvgo_mailbox_export(h, buf), 'islam_pvgo_mailbox_export')
            mine = torch.tensor(list(buf), dtype=torch.uint8, device=self.s.device if self._nccl else 'cpu')
            table = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(table, mine, group=group)
            raw = torch.stack(table).cpu().numpy().tobytes()
            _lib.check(L.islam_pvgo_mailbox_connect(h, raw), 'islam_pvgo_mailbox_connect')

    # passthroughs
    def set_problem(self, *a, **k):
        self.s.set_problem(*a, **k)

    def set_state(self, nodes, vels):
        self.s.set_state(nodes, vels)

    def lm_reset(self, **kw):
        self.s.lm_reset(**kw)

    def _allreduce(self, t):
        if self._nccl:
            dist.all_reduce(t, group=self.group)
        else:                                   # gloo: stage through the host
            c = t.cpu()
            dist.all_reduce(c, group=self.group)
            t.copy_(c)

    def _bcast(self, t, src):
        if self.root_exchange == 'allreduce':   # owner's block + zeros everywhere else (islam_pvgo_root_zero_foreign)
            return self._allreduce(t)
        if self._nccl:
            dist.broadcast(t, src, group=self.group)
        else:                                   # gloo: stage through the host
            c = t.cpu()
            dist.broadcast(c, src, group=self.group)
            t.copy_(c)

    def _root_factor(self, st):
        """Dense root, 1-D block-column-cyclic right-looking Cholesky (include/islam_pvgo.h): per 128-column block the owner
        factors it, broadcasts it (NCCL over NVSwitch), and every rank updates its own tile columns of the trailing matrix
        with it (the n^3/3 of the work, split G ways).  Look-ahead: the NEXT block's tile column is updated first, so that
        its owner factors it and the broadcast travels on a second stream while all ranks are still busy with the rest of
        this block's update — the chain of panel factorisations and broadcasts leaves the critical path."""
        s, n, ld, nb = self.s, self.root_n, self.root_ld, self.root_block
        L, h = s.L, s._h
        main = s.stream
        if self._side is None:
            # high priority: the panel kernels and the broadcast must get SMs while the wide update of the previous block
            # still has thousands of CTAs queued, or the look-ahead only starts when that update drains.  For the same
            # reason launch with TORCH_NCCL_HIGH_PRIORITY=1 (torch then creates its NCCL streams with high priority).
            self._side = torch.cuda.Stream(device=s.device, priority=-1)
            self._root_ev = [torch.cuda.Event() for _ in range(2)]
        side, side_p = self._side, C.c_void_p(self._side.cuda_stream)
        blocks = list(range(0, n, nb))
        if self.root_exchange == 'allreduce':
            _lib.check(L.islam_pvgo_root_zero_foreign(h, st), 'islam_pvgo_root_zero_foreign')
        _lib.check(L.islam_pvgo_root_panel(h, 0, st), 'islam_pvgo_root_panel')
        self._bcast(self.root_R[0:min(nb, n) * ld], self._root_src[0])
        for b, k0 in enumerate(blocks):
            k1 = k0 + nb
            if not self.lookahead or k1 >= n:
                _lib.check(L.islam_pvgo_root_update(h, k0, st), 'islam_pvgo_root_update')
                if k1 < n:
                    _lib.check(L.islam_pvgo_root_panel(h, k1, st), 'islam_pvgo_root_panel')
                    self._bcast(self.root_R[k1 * ld:min(k1 + nb, n) * ld], self._root_src[b + 1])
                continue
            _lib.check(L.islam_pvgo_root_update_part(h, k0, 1, st), 'islam_pvgo_root_update_part')
            ready, done = self._root_ev
            ready.record(main)
            side.wait_event(ready)
            with torch.cuda.stream(side):
                _lib.check(L.islam_pvgo_root_panel(h, k1, side_p), 'islam_pvgo_root_panel')
                self._bcast(self.root_R[k1 * ld:min(k1 + nb, n) * ld], self._root_src[b + 1])
                done.record(side)
            _lib.check(L.islam_pvgo_root_update_part(h, k0, 2, st), 'islam_pvgo_root_update_part')
            main.wait_event(done)

    _side = None
    lookahead = os.environ.get('ISLAM_ROOT_LOOKAHEAD', '1') != '0'
    # how a factored block reaches the other ranks: 'broadcast' (NCCL broadcast from the owner) or 'allreduce' (SUM of the
    # owner's block and zeros: in-fabric reduction + multicast on NVSwitch)
    root_exchange = os.environ.get('ISLAM_ROOT_EXCHANGE', 'broadcast')

    def lm_try(self):
        s = self.s
        st = C.c_void_p(s.stream.cuda_stream)
        with torch.cuda.stream(s.stream):
            _lib.check(s.L.islam_pvgo_lm_try_begin(s._h, st), 'islam_pvgo_lm_try_begin')
            self._allreduce(self.shared)
            if self.root_n:
                self._allreduce(self.root_R)
                self._allreduce(self.root_diag)
            _lib.check(s.L.islam_pvgo_lm_try_mid(s._h, st), 'islam_pvgo_lm_try_mid')
            if self.root_n:
                self._root_factor(st)
                _lib.check(s.L.islam_pvgo_lm_try_mid2(s._h, st), 'islam_pvgo_lm_try_mid2')
            if self.exchange == 'nccl':
                self._allreduce(self.sums)
            _lib.check(s.L.islam_pvgo_lm_try_end(s._h, st), 'islam_pvgo_lm_try_end')

    def _graph_try(self):
        """One try as a CUDA graph: ~25 kernels of the library (with their programmatic dependent launches) AND the NCCL
        all-reduce between them, captured once on the solver's stream and replayed — the single-GPU path has always
        replayed a graph (islam_pvgo_lm_run); the sharded one used to enqueue every kernel from Python through three ctypes
        calls around torch.distributed.  All control state lives on the device, so the captured arguments never change.
        Falls back to eager enqueueing with the gloo test rig (host-staged collectives cannot be captured)."""
        if not self._nccl or self._graph is False:
            return None
        if self._graph is None:
            try:
                self.lm_try()                                   # warm-up outside capture (NCCL channel set-up, lazy allocations)
                torch.cuda.synchronize(self.s.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.s.stream, capture_error_mode='thread_local'):
                    self.lm_try()
                self._graph = g
            except Exception:                                   # capture not possible in this environment: stay eager
                self._graph = False
                return None
        return self._graph

    _graph = None

    def lm_run(self, budget=None):
        """The `while scheduler.continual()` loop: tries are enqueued back to back (device-side predicates make
        surplus tries no-ops); one synchronisation at the end, more only if rejected tries exhaust the budget."""
        s = self.s
        s._enter()
        if budget is None:                      # same policy as islam_pvgo_lm_run
            budget = min(s.params.max_steps + 2, 4) if s.params.use_scheduler else s.params.max_steps + 2
        if self.root_n:                         # a surplus try would still broadcast the whole root: check after every try
            budget = 1
        g = self._graph_try() if self.use_graph and not self.root_n else None
        # worst case: every step burns its 16 rejected tries before it is accepted or abandoned (PyPose's reject=16)
        for _ in range(2 + 17 * max(1, s.params.max_steps) // (1 if self.root_n else 4)):
            for _ in range(budget):
                if g is not None:
                    with torch.cuda.stream(s.stream):
                        g.replay()
                else:
                    self.lm_try()
            st = s.lm_state()
            if not st.continual:
                return st
            budget = 1 if self.root_n else 4
        # as islam_pvgo_lm_run's -9: the loop did not close within its worst-case try budget
        raise IslamError('ShardedPVGO.lm_run: the LM loop did not finish within its try budget (max_steps too large?)')

    # opt-in (ISLAM_SHARDED_GRAPH=1): replaying the captured try measured 2 271 LM it/s against 2 254 eager on two B200 (the
    # critical path is the chain of tree levels, not the launches), and tearing the process group down while a graph that
    # captured NCCL kernels is alive hung the workers at exit — call release_graph() before destroy_process_group()
    use_graph = os.environ.get('ISLAM_SHARDED_GRAPH') == '1'

    def release_graph(self):
        self._graph = None

    def get_state(self):
        """Each pose is taken from the rank that solves it (shared poses from rank 0): ONE all-reduce of the packed
        (N, 10) state whose foreign rows are zeroed."""
        n, v = self.s.get_state()
        packed = torch.cat([torch.where(self._mine_pose.unsqueeze(-1), n, torch.zeros_like(n)),
                            torch.where(self._mine_vel.unsqueeze(-1), v, torch.zeros_like(v))], dim=1)
        self._allreduce(packed)
        return packed[:, :7].contiguous(), packed[:, 7:].contiguous()
