"""pp.optim.LM for PoseVelGraph-shaped models (see package docstring).  Restates the attribute surface of PyPose's
LevenbergMarquardt (loss, last, reject, reject_count, step(input, weight)) used at /root/reference/pvgo.py:169-180."""
import numpy as np
import torch

from . import solver, strategy, scheduler, kernel, corrector      # noqa: F401
from ..._lib import IslamError


def _scalar_info(w, name):
    """The reference's information matrices are c * I (pvgo.py:125-143); recover c and refuse anything else."""
    w = w.detach().as_subclass(torch.Tensor)
    c = float(w.reshape(-1, w.shape[-2], w.shape[-1])[0, 0, 0])
    eye = torch.eye(w.shape[-1], device=w.device, dtype=w.dtype) * c
    if not torch.allclose(w, eye.expand_as(w), rtol=1e-6, atol=0):
        raise NotImplementedError(f'{name}: only scalar-diagonal information matrices (as built by run_pvgo) are supported')
    return c


class LM:
    def __init__(self, model, solver=None, strategy=None, kernel=None, corrector=None, weight=None, reject=16,
                 min=1e-6, max=1e32, vectorize=True):
        if not (hasattr(model, 'nodes') and hasattr(model, 'vels')):
            raise NotImplementedError('pp.optim.LM here is the fused PVGO path: the model must expose `nodes` (N,7) and '
                                      '`vels` (N,3) like pvgo.PoseVelGraph; there is no generic autograd LM')
        if getattr(model, 'reproj', None) is not None:
            raise NotImplementedError('the optional reprojection factor (pvgo.py:53-61) is not on the B200 path yet')
        self.model, self.strategy = model, strategy
        self.reject, self.min, self.max = reject, min, max
        self.reject_count = 0
        self.last = None
        self._solver = None
        self._key = None

    def _bind(self, input, weight):
        from ...solver import PVGOSolver
        edges, poses, drots, dtrans, dvels, dts = input
        nodes = self.model.nodes.detach().as_subclass(torch.Tensor)
        dev = nodes.device
        if dev.type != 'cuda':
            raise IslamError('pp.optim.LM needs the model on a CUDA device (no CPU fallback)')
        e = edges.detach().cpu().numpy().astype(np.int64)
        key = (nodes.shape[0], e.tobytes())
        if self._solver is None or self._key != key:
            self._solver = PVGOSolver(nodes.shape[0], e, device=dev)
            self._key = key
            self._solver.set_state(nodes, self.model.vels.detach())
            radius = getattr(self.strategy, 'radius', 1e6) if self.strategy is not None else 1e6
            self._solver.lm_reset(radius=float(radius), lm_min=float(self.min), lm_max=float(self.max),
                                  reject=int(self.reject), max_steps=1 << 30, use_scheduler=0)
        if weight is None:
            w = (1.0, 1.0, 1.0, 1.0)
        else:
            if len(weight) != 4:
                raise NotImplementedError('expected the 4 weight groups of pvgo.py:162')
            w = tuple(_scalar_info(x, f'weight[{i}]') for i, x in enumerate(weight))
        lw = tuple(np.sqrt(x) for x in w)                 # the solver squares loss weights (pvgo.py:125-129)
        self._solver.set_problem(poses, drots, dtrans, dvels, dts.reshape(-1), lw)

    @torch.no_grad()
    def step(self, input, target=None, weight=None):
        self._bind(input, weight)
        st = self._solver.lm_step()
        n, v = self._solver.get_state()
        self.model.nodes.as_subclass(torch.Tensor).copy_(n)
        self.model.vels.copy_(v)
        self.loss = torch.tensor(st.loss, device=n.device, dtype=torch.float32)
        self.last = torch.tensor(st.last, device=n.device, dtype=torch.float32)
        self.reject_count = st.reject_count
        if st.info:
            print('Linear solver failed. Breaking optimization step...')
        return self.loss


LevenbergMarquardt = LM
