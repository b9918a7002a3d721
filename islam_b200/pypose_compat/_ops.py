"""LieTensor arithmetic for the PyPose-compatible shim: every map runs in the CUDA kernels of csrc/lieops.cu through
the C ABI (islam_lie_*).  Tensors that live on the host (train.py keeps window poses on the CPU, e.g.
/root/reference/train.py:239-240) are staged through cuda:0 and copied back — there is no CPU arithmetic path.

Autograd follows PyPose's convention: the gradient of a group-valued tensor is the left-tangent
gradient stored in the leading slots of an embedding-sized row."""
import ctypes as C

import torch

from .. import _lib
from .._lib import IslamError

SE3, SO3 = 0, 1
_EMB = {SE3: 7, SO3: 4}
_ALG = {SE3: 6, SO3: 3}


def _dev(t):
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise IslamError('LieTensor arithmetic runs in CUDA kernels only (no CPU fallback) and no CUDA device is visible')
    return torch.device('cuda', torch.cuda.current_device())


def _prep(t, dev, width):
    t = t.detach().as_subclass(torch.Tensor) if type(t) is not torch.Tensor else t.detach()
    if t.shape[-1] != width:
        raise IslamError(f'expected last dimension {width}, got {tuple(t.shape)}')
    return t.to(device=dev, dtype=torch.float32).contiguous()


def _call(name, group, ins, out_shapes, dev):
    L = _lib.lib()
    outs = [torch.empty(s, dtype=torch.float32, device=dev) if s is not None else None for s in out_shapes]
    n = 1
    for d in ins[0].shape[:-1]:
        n *= d
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(getattr(L, name)(group, *[p(t) for t in ins], *[p(t) for t in outs], n, stream), name)
    return outs


def _back(t, like):
    return t.to(device=like.device, dtype=like.dtype if like.dtype.is_floating_point else torch.float32)


def _bcast(a, b, wa, wb):
    """Broadcast the batch dimensions of a (..., wa) and b (..., wb)."""
    sa, sb = a.shape[:-1], b.shape[:-1]
    shape = torch.broadcast_shapes(sa, sb)
    return a.expand(*shape, wa), b.expand(*shape, wb)


class ExpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        dev = _dev(x)
        xd = _prep(x, dev, _ALG[group])
        (y,) = _call('islam_lie_exp', group, [xd], [xd.shape[:-1] + (_EMB[group],)], dev)
        ctx.save_for_backward(xd)
        ctx.group, ctx.like = group, x
        return _back(y, x)

    @staticmethod
    def backward(ctx, gy):
        (xd,) = ctx.saved_tensors
        g = _prep(gy, xd.device, _EMB[ctx.group])
        (gx,) = _call('islam_lie_exp_bwd', ctx.group, [xd, g], [xd.shape], xd.device)
        return _back(gx, gy), None


class LogFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        dev = _dev(x)
        xd = _prep(x, dev, _EMB[group])
        (y,) = _call('islam_lie_log', group, [xd], [xd.shape[:-1] + (_ALG[group],)], dev)
        ctx.save_for_backward(y)
        ctx.group = group
        return _back(y, x)

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        g = _prep(gy, y.device, _ALG[ctx.group])
        (gx,) = _call('islam_lie_log_bwd', ctx.group, [y, g], [y.shape[:-1] + (_EMB[ctx.group],)], y.device)
        return _back(gx, gy), None


class InvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        dev = _dev(x)
        xd = _prep(x, dev, _EMB[group])
        (y,) = _call('islam_lie_inv', group, [xd], [xd.shape], dev)
        ctx.save_for_backward(y)
        ctx.group = group
        return _back(y, x)

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        g = _prep(gy, y.device, _EMB[ctx.group])
        (gx,) = _call('islam_lie_inv_bwd', ctx.group, [y, g], [y.shape], y.device)
        return _back(gx, gy), None


class MulFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, group):
        dev = _dev(a if a.is_cuda or not b.is_cuda else b)
        w = _EMB[group]
        ae, be = _bcast(a, b, w, w)
        ad, bd = _prep(ae, dev, w), _prep(be, dev, w)
        (y,) = _call('islam_lie_mul', group, [ad, bd], [ad.shape], dev)
        ctx.save_for_backward(ad)
        ctx.group, ctx.sa, ctx.sb = group, a.shape, b.shape
        return _back(y, a)

    @staticmethod
    def backward(ctx, gy):
        (ad,) = ctx.saved_tensors
        g = _prep(gy.expand(ad.shape), ad.device, _EMB[ctx.group])
        ga, gb = _call('islam_lie_mul_bwd', ctx.group, [ad, g], [ad.shape, ad.shape], ad.device)
        return _back(ga, gy).sum_to_size(ctx.sa), _back(gb, gy).sum_to_size(ctx.sb), None


class ActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, group):
        dev = _dev(x if x.is_cuda or not p.is_cuda else p)
        xe, pe = _bcast(x, p, _EMB[group], 3)
        xd, pd = _prep(xe, dev, _EMB[group]), _prep(pe, dev, 3)
        (y,) = _call('islam_lie_act', group, [xd, pd], [pd.shape], dev)
        ctx.save_for_backward(xd, pd)
        ctx.group, ctx.sx, ctx.sp = group, x.shape, p.shape
        return _back(y, p)

    @staticmethod
    def backward(ctx, gy):
        xd, pd = ctx.saved_tensors
        g = _prep(gy.expand(pd.shape), xd.device, 3)
        gx, gp = _call('islam_lie_act_bwd', ctx.group, [xd, pd, g], [xd.shape, pd.shape], xd.device)
        return _back(gx, gy).sum_to_size(ctx.sx), _back(gp, gy).sum_to_size(ctx.sp), None


def cumprod(x, group, left=True):
    """Ordered prefix product along dim 0 of a (n, W) tensor; no autograd (the reference only uses it detached)."""
    dev = _dev(x)
    xd = _prep(x, dev, _EMB[group])
    if xd.dim() != 2:
        raise IslamError('cumprod expects a (n, W) LieTensor')
    y = torch.empty_like(xd)
    L = _lib.lib()
    with torch.cuda.device(dev):
        _lib.check(L.islam_lie_cumprod(group, C.c_void_p(xd.data_ptr()), C.c_void_p(y.data_ptr()), xd.shape[0],
                                       1 if left else 0, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                   'islam_lie_cumprod')
    return _back(y, x)
