"""pp.optim.LM for PoseVelGraph-shaped models (see package docstring).  Restates the attribute surface of PyPose's
LevenbergMarquardt (loss, last, reject, reject_count, step(input, weight)) used at /root/reference/pvgo.py:169-180."""
import numpy as np
import torch

import sys
import types

from . import strategy, scheduler      # noqa: F401
from ..._lib import IslamError


class Cholesky:
    """pypose.optim.solver.Cholesky (pvgo.py:169): a configuration token — the linear algebra is the multifrontal float64
    Cholesky of csrc/front4.cuh / solver3.cuh."""

    def __init__(self, upper=False):
        self.upper = upper


class Trivial:
    """pypose.optim.kernel.Trivial: the identity robust kernel (what LM uses when kernel=None, pvgo.py:171)."""


def _unsupported(what, where):
    class _Unsupported:
        def __init__(self, *a, **kw):
            raise NotImplementedError(f'{what} is not used by iSLAM ({where}) and is not provided on the B200 path')
    _Unsupported.__name__ = what
    return _Unsupported


def _namespace(name, **attrs):
    """pvgo.py:8-10 imports pypose.optim.{solver,kernel,corrector} as modules; they only carry these names."""
    m = types.ModuleType(f'{__name__}.{name}')
    m.__dict__.update(attrs)
    sys.modules[m.__name__] = m
    return m


solver = _namespace('solver', Cholesky=Cholesky, PINV=_unsupported('PINV', 'pvgo.py:169 uses Cholesky'),
                    LSTSQ=_unsupported('LSTSQ', 'pvgo.py:169 uses Cholesky'))
kernel = _namespace('kernel', Trivial=Trivial, Huber=_unsupported('Huber', 'pvgo.py:171 passes no kernel'),
                    PseudoHuber=_unsupported('PseudoHuber', 'pvgo.py:171 passes no kernel'),
                    Cauchy=_unsupported('Cauchy', 'pvgo.py:171 passes no kernel'))
corrector = _namespace('corrector', FastTriggs=_unsupported('FastTriggs', 'pvgo.py:171 passes no corrector'),
                       Triggs=_unsupported('Triggs', 'pvgo.py:171 passes no corrector'))


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
        self.model, self.strategy = model, strategy
        self.reject, self.min, self.max = reject, min, max
        self.reject_count = 0
        self.last = self.loss = None
        self._solver = None

    def __del__(self):
        s, self._solver = getattr(self, '_solver', None), None
        if s is not None:
            try:
                from ... import pvgo as _pvgo
                _pvgo.release_solver(s)
            except Exception:      # interpreter shutdown
                pass

    def _bind(self, input, weight):
        from ... import pvgo as _pvgo
        edges, poses, drots, dtrans, dvels, dts = input
        nodes = self.model.nodes.detach().as_subclass(torch.Tensor)
        dev = nodes.device
        if dev.type != 'cuda':
            raise IslamError('pp.optim.LM needs the model on a CUDA device (no CPU fallback)')
        if self._edges_obj is not edges or self._solver is None:
            e = edges.detach().cpu().numpy().astype(np.int64).reshape(-1, 2)
            if self._solver is None or self._solver.N != nodes.shape[0] or not np.array_equal(self._solver.links, e):
                if self._solver is not None:
                    _pvgo.release_solver(self._solver)
                # the symbolic analysis of a graph structure is reused across optimiser objects (same pool as run_pvgo)
                self._solver = _pvgo.acquire_solver(nodes.shape[0], e, dev)
                self._fresh = True
            self._edges_obj = edges
        if self._fresh:
            self._fresh = False
            self._solver.set_state(nodes, self.model.vels.detach())
            kw = self.strategy.lm_params() if hasattr(self.strategy, 'lm_params') else {}
            if self.strategy is None:
                kw = dict(radius=1e6)
            self._solver.lm_reset(lm_min=float(self.min), lm_max=float(self.max), reject=int(self.reject),
                                  max_steps=1 << 30, use_scheduler=0, **kw)
        reproj = getattr(self.model, 'reproj', None)
        n_w = 4 + (reproj is not None)
        if weight is None:
            w = (1.0,) * n_w
        else:
            if len(weight) != n_w:
                raise NotImplementedError(f'expected the {n_w} weight groups of pvgo.py:162-165')
            w = tuple(_scalar_info(x, f'weight[{i}]') for i, x in enumerate(weight))
        lw = [float(np.sqrt(x)) for x in w]               # the solver squares loss weights (pvgo.py:125-129)
        if reproj is not None:
            lw[4] *= reproj.N                              # pvgo.py:131: info = (loss_weight[4] / N)^2
        self._solver.set_problem(poses, drots, dtrans, dvels, dts.reshape(-1), tuple(lw), reproj=reproj)

    _edges_obj = None
    _fresh = True

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
        if st.info == 1:
            print('Linear solver failed. Breaking optimization step...')
        elif st.info:
            raise IslamError(f'the device-side LM step reported info={st.info} (see include/islam_pvgo.h)')
        return self.loss


LevenbergMarquardt = LM
