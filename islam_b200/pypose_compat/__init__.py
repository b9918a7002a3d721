"""PyPose-compatible surface for the names iSLAM's back-end is written against, backed by the
B200 kernels.  `install()` registers this package as `pypose` in sys.modules so that the reference's own
pvgo.py / imu_integrator.py / Datasets/transformation.py import and run unchanged:

    import islam_b200.pypose_compat as ppc; ppc.install()
    import pypose as pp            # -> this module

Covered: pp.SE3 / SO3 / se3 / so3 / LieTensor / Parameter / identity_SO3 / identity_SE3 / from_matrix / *_type,
LieTensor.{Exp, Log, Inv, rotation, translation, tensor, matrix, lview, ltype}, `@` / `*` (group x group, group x points),
shape ops that keep the ltype (indexing, stack, cat, to, cpu, clone, detach, ...), pp.module.IMUPreintegrator,
pp.optim.LM (+ .solver.Cholesky, .strategy.TrustRegion, .scheduler.StopOnPlateau, .kernel, .corrector) — the optimiser is
NOT a generic autograd LM: it pattern-matches PoseVelGraph-shaped models (attributes `nodes`, `vels`, 6-tuple input,
/root/reference/pvgo.py:15-64) onto the fused CUDA path and raises for anything else.
"""
import sys
import types

import numpy as np
import torch
from torch.utils._pytree import tree_flatten, tree_map

from . import _ops
from .._lib import IslamError


class LieType:
    def __init__(self, name, group, dimension, manifold, is_group):
        self.name, self.group, self.dimension, self.manifold, self.is_group = name, group, dimension, manifold, is_group

    def __repr__(self):
        return self.name


SE3_type = LieType('SE3_type', _ops.SE3, 7, 6, True)
se3_type = LieType('se3_type', _ops.SE3, 6, 6, False)
SO3_type = LieType('SO3_type', _ops.SO3, 4, 3, True)
so3_type = LieType('so3_type', _ops.SO3, 3, 3, False)
_PARTNER = {SE3_type: se3_type, se3_type: SE3_type, SO3_type: so3_type, so3_type: SO3_type}

# torch functions whose tensor outputs keep the LieTensor type (same list idea as PyPose's HANDLED_FUNCTIONS)
_KEEP = {'__getitem__', '__setitem__', 'cpu', 'cuda', 'float', 'double', 'to', 'detach', 'view', 'view_as', 'squeeze',
         'unsqueeze', 'cat', 'concat', 'concatenate', 'stack', 'split', 'chunk', 'tensor_split', 'index_select',
         'masked_select', 'movedim', 'moveaxis', 'narrow', 'permute', 'reshape', 'clone', 'swapaxes', 'swapdims',
         'take_along_dim', 'tile', 'transpose', 'unbind', 'gather', 'repeat', 'expand', 'expand_as', 'select',
         'index_put', 'index_put_', 'copy_', 'contiguous', 'flip', 'roll', 'requires_grad_', 'pin_memory',
         'vstack', 'hstack', 'row_stack'}


class LieTensor(torch.Tensor):
    """A torch.Tensor whose last dimension stores a Lie group / algebra element."""

    @staticmethod
    def __new__(cls, data, ltype=None, **kw):
        if isinstance(data, LieTensor) and ltype is None:
            ltype = data.ltype
        if not isinstance(data, torch.Tensor):
            data = torch.as_tensor(np.asarray(data, dtype=np.float32) if not isinstance(data, np.ndarray) else data)
            if data.dtype == torch.float64 and torch.get_default_dtype() == torch.float32:
                data = data.float()
        if ltype is None:
            raise IslamError('LieTensor needs an ltype')
        if data.shape[-1] != ltype.dimension:
            raise IslamError(f'{ltype} expects last dimension {ltype.dimension}, got {tuple(data.shape)}')
        base = data.as_subclass(torch.Tensor) if type(data) is not torch.Tensor else data
        t = base.as_subclass(cls)
        t.ltype = ltype
        return t

    def __init__(self, *a, **kw):
        pass

    @classmethod
    def __torch_function__(cls, func, types_, args=(), kwargs=None):
        kwargs = kwargs or {}
        flat, _ = tree_flatten((args, kwargs))
        lt = next((a.ltype for a in flat if isinstance(a, LieTensor) and getattr(a, 'ltype', None) is not None), None)
        with torch._C.DisableTorchFunctionSubclass():
            out = func(*args, **kwargs)
        name = getattr(func, '__name__', '')
        keep = name in _KEEP and lt is not None

        def fix(t):
            if isinstance(t, torch.Tensor):
                if keep and t.dim() > 0 and t.shape[-1] == lt.dimension and t.dtype.is_floating_point:
                    if not isinstance(t, LieTensor):
                        t = t.as_subclass(LieTensor)
                    t.ltype = lt
                    return t
                if isinstance(t, LieTensor):
                    return t.as_subclass(torch.Tensor)
            return t
        return tree_map(fix, out)

    # ------------------------------------------------------------------------------------------ views
    def tensor(self):
        return self.as_subclass(torch.Tensor)

    def __repr__(self):
        return f'{self.ltype.name[:-5]}LieTensor:\n' + repr(self.tensor())

    def lview(self, *shape):
        return LieTensor(self.tensor().reshape(*shape, self.ltype.dimension), ltype=self.ltype)

    @property
    def lshape(self):
        return self.shape[:-1]

    def translation(self):
        assert self.ltype is SE3_type
        return self.tensor()[..., :3]

    def rotation(self):
        if self.ltype is SO3_type:
            return self
        assert self.ltype is SE3_type
        return LieTensor(self.tensor()[..., 3:7], ltype=SO3_type)

    def matrix(self):
        t = self.tensor()
        if self.ltype is SE3_type:
            R = _quat_to_matrix(t[..., 3:7])
            top = torch.cat([R, t[..., :3].unsqueeze(-1)], -1)
            bot = torch.zeros_like(top[..., :1, :])
            bot[..., 0, 3] = 1
            return torch.cat([top, bot], -2)
        if self.ltype is SO3_type:
            return _quat_to_matrix(t)
        raise IslamError('matrix() needs a group element')

    # ------------------------------------------------------------------------------------------ maps
    def Exp(self):
        assert not self.ltype.is_group, 'Exp maps an algebra element'
        return LieTensor(_ops.ExpFn.apply(self.tensor(), self.ltype.group), ltype=_PARTNER[self.ltype])

    def Log(self):
        assert self.ltype.is_group, 'Log maps a group element'
        return LieTensor(_ops.LogFn.apply(self.tensor(), self.ltype.group), ltype=_PARTNER[self.ltype])

    def Inv(self):
        if not self.ltype.is_group:
            return LieTensor(-self.tensor(), ltype=self.ltype)
        return LieTensor(_ops.InvFn.apply(self.tensor(), self.ltype.group), ltype=self.ltype)

    def Act(self, p):
        return _ops.ActFn.apply(self.tensor(), _plain(p), self.ltype.group)

    def __matmul__(self, other):
        assert self.ltype.is_group, '@ needs a group element on the left'
        if isinstance(other, LieTensor):
            if other.ltype is not self.ltype:
                raise IslamError(f'cannot compose {self.ltype} with {other.ltype}')
            return LieTensor(_ops.MulFn.apply(self.tensor(), other.tensor(), self.ltype.group), ltype=self.ltype)
        other = _plain(other)
        if other.shape[-1] == 3:
            return _ops.ActFn.apply(self.tensor(), other, self.ltype.group)
        raise IslamError('group @ tensor expects (..., 3) points')

    def __mul__(self, other):
        if isinstance(other, LieTensor) or (isinstance(other, torch.Tensor) and other.dim() > 0 and other.shape[-1] == 3
                                            and self.ltype.is_group):
            return self.__matmul__(other)
        return LieTensor(self.tensor() * other, ltype=self.ltype) if not self.ltype.is_group else NotImplemented

    def add_(self, other):
        """LieTensor.add_: X <- Exp(other[..., :6]) * X for groups (A.1), plain addition for algebras."""
        o = _plain(other)
        if self.ltype.is_group:
            d = LieTensor(o[..., :self.ltype.manifold], ltype=_PARTNER[self.ltype])
            new = (d.Exp() @ LieTensor(self.tensor().detach(), ltype=self.ltype)).tensor()
            with torch.no_grad():
                self.tensor().copy_(new)
            return self
        with torch.no_grad():
            self.tensor().add_(o[..., :self.ltype.dimension])
        return self

    def numpy(self):
        return self.tensor().detach().numpy() if not self.requires_grad else self.tensor().numpy()


class Parameter(LieTensor, torch.nn.Parameter):
    """pp.Parameter: a LieTensor that is an nn.Parameter (pvgo.py:20)."""

    @staticmethod
    def __new__(cls, data, requires_grad=True):
        if not isinstance(data, LieTensor):
            raise IslamError('pp.Parameter wraps a LieTensor')
        t = torch.Tensor._make_subclass(cls, data.tensor().detach(), requires_grad)
        t.ltype = data.ltype
        return t

    def __deepcopy__(self, memo):
        r = Parameter(LieTensor(self.tensor().detach().clone(), ltype=self.ltype), self.requires_grad)
        memo[id(self)] = r
        return r


def _plain(t):
    if isinstance(t, torch.Tensor):
        return t.as_subclass(torch.Tensor) if type(t) is not torch.Tensor else t
    return torch.as_tensor(np.asarray(t, dtype=np.float32))


def _quat_to_matrix(q):
    x, y, z, w = q.unbind(-1)
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        torch.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)


def _matrix_to_quat(R):
    """Rotation matrix -> (x,y,z,w), branch on the largest diagonal term (constant frames only; not a hot path)."""
    R = R.to(torch.float64)
    m = lambda i, j: R[..., i, j]
    tr = m(0, 0) + m(1, 1) + m(2, 2)
    cands = torch.stack([
        torch.stack([m(2, 1) - m(1, 2), m(0, 2) - m(2, 0), m(1, 0) - m(0, 1), 1 + tr], -1),
        torch.stack([1 + m(0, 0) - m(1, 1) - m(2, 2), m(0, 1) + m(1, 0), m(0, 2) + m(2, 0), m(2, 1) - m(1, 2)], -1),
        torch.stack([m(0, 1) + m(1, 0), 1 - m(0, 0) + m(1, 1) - m(2, 2), m(1, 2) + m(2, 1), m(0, 2) - m(2, 0)], -1),
        torch.stack([m(0, 2) + m(2, 0), m(1, 2) + m(2, 1), 1 - m(0, 0) - m(1, 1) + m(2, 2), m(1, 0) - m(0, 1)], -1)], -2)
    diag = torch.stack([tr, m(0, 0), m(1, 1), m(2, 2)], -1)
    idx = diag.argmax(-1)
    q = torch.gather(cands, -2, idx[..., None, None].expand(*idx.shape, 1, 4)).squeeze(-2)
    q = q / q.norm(dim=-1, keepdim=True)
    return q


def SE3(data, **kw):
    return LieTensor(data, ltype=SE3_type)


def SO3(data, **kw):
    return LieTensor(data, ltype=SO3_type)


def se3(data, **kw):
    return LieTensor(data, ltype=se3_type)


def so3(data, **kw):
    return LieTensor(data, ltype=so3_type)


def identity_SO3(*lsize, **kw):
    t = torch.zeros(*lsize, 4, **kw)
    t[..., 3] = 1
    return LieTensor(t, ltype=SO3_type)


def identity_SE3(*lsize, **kw):
    t = torch.zeros(*lsize, 7, **kw)
    t[..., 6] = 1
    return LieTensor(t, ltype=SE3_type)


def from_matrix(mat, ltype=SE3_type, **kw):
    mat = torch.as_tensor(np.asarray(mat, dtype=np.float64)) if not isinstance(mat, torch.Tensor) else mat
    if ltype is SE3_type:
        q = _matrix_to_quat(mat[..., :3, :3])
        return LieTensor(torch.cat([mat[..., :3, 3].to(torch.float64), q], -1).to(torch.get_default_dtype()), ltype=SE3_type)
    if ltype is SO3_type:
        return LieTensor(_matrix_to_quat(mat[..., :3, :3]).to(torch.get_default_dtype()), ltype=SO3_type)
    raise IslamError('from_matrix supports SE3_type / SO3_type')


def mat2SE3(mat, **kw):
    return from_matrix(mat, SE3_type)


def mat2SO3(mat, **kw):
    return from_matrix(mat, SO3_type)


def Exp(x):
    return x.Exp()


def Log(x):
    return x.Log()


def Inv(x):
    return x.Inv()


def cumprod(x, dim=0, left=True):
    """pp.cumprod along the leading dimension (prefix-product scan kernel, csrc/lieops.cu)."""
    if dim not in (0, -2) or x.dim() != 2:
        raise IslamError('cumprod is implemented along the leading dimension of a (n, W) LieTensor')
    return LieTensor(_ops.cumprod(x.tensor(), x.ltype.group, left), ltype=x.ltype)


from . import module, optim, function          # noqa: E402
from .optim import solver as _solver, strategy as _strategy, scheduler as _scheduler, kernel as _kernel, corrector as _corrector  # noqa: E402,F401


def install(name='pypose', force=False):
    """Register this package (and its sub-modules) under `name` so `import pypose as pp` resolves here."""
    if name in sys.modules and not force and sys.modules[name] is not sys.modules[__name__]:
        real = sys.modules[name]
        if getattr(real, '__islam_shim__', False) is False:
            raise IslamError(f'a real `{name}` package is already imported; refusing to shadow it')
    me = sys.modules[__name__]
    me.__islam_shim__ = True
    sys.modules[name] = me
    for sub in ('module', 'optim', 'function', 'optim.solver', 'optim.strategy', 'optim.scheduler', 'optim.kernel',
                'optim.corrector', 'function.geometry'):
        sys.modules[f'{name}.{sub}'] = sys.modules[f'{__name__}.{sub}']
    return me
