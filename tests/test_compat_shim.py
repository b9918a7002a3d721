"""Host logic of the PyPose-compatible shim that needs no arithmetic: type plumbing, module registration, scheduler."""
import sys

import numpy as np
import pytest
import torch

import islam_b200.pypose_compat as ppc


def test_install_registers_pypose_names():
    ppc.install()
    import pypose as pp
    import pypose.optim.solver as ppos
    import pypose.optim.kernel as ppok          # noqa: F401  (pvgo.py:9)
    import pypose.optim.corrector as ppoc       # noqa: F401  (pvgo.py:10)
    import pypose.optim.strategy as ppost
    from pypose.optim.scheduler import StopOnPlateau
    from pypose.function.geometry import reprojerr, point2pixel      # noqa: F401  (dense_ba.py:5)
    assert pp is ppc and ppos.Cholesky and ppost.TrustRegion(radius=1e4).radius == 1e4 and StopOnPlateau
    for name in ('SE3', 'SO3', 'se3', 'so3', 'LieTensor', 'SE3_type', 'se3_type', 'Parameter', 'identity_SO3',
                 'from_matrix', 'module', 'optim'):
        assert hasattr(pp, name), name
    assert hasattr(pp.module, 'IMUPreintegrator') and hasattr(pp.optim, 'LM')


def test_ltype_survives_shape_ops():
    x = ppc.SE3(torch.arange(21, dtype=torch.float32).reshape(3, 7))
    assert x.ltype is ppc.SE3_type and x[0].ltype is ppc.SE3_type and x[1:].ltype is ppc.SE3_type
    assert x[torch.tensor([0, 2])].shape == (2, 7)
    assert torch.stack([x[0], x[1]]).ltype is ppc.SE3_type                   # transformation.py:113
    assert torch.cat([x, x]).ltype is ppc.SE3_type
    for y in (x.clone(), x.detach(), x.cpu(), x.to(torch.float64), x.unsqueeze(0), x.unsqueeze(0).squeeze(0)):
        assert isinstance(y, ppc.LieTensor) and y.ltype is ppc.SE3_type
    assert type(x.tensor()) is torch.Tensor and type(x.translation()) is torch.Tensor
    assert x.rotation().ltype is ppc.SO3_type and x.rotation().shape == (3, 4)
    assert type(x.sum()) is torch.Tensor and type(x[..., :3]) is torch.Tensor   # no longer a Lie element
    assert isinstance(x.numpy(), np.ndarray)
    y = x.clone()
    y[0] = 0.1                                                                # pvgo.py:57 __setitem__
    assert float(y[0, 3]) == pytest.approx(0.1)
    assert x.lview(3, 1).shape == (3, 1, 7)


def test_parameter_is_an_nn_parameter_with_ltype():
    p = ppc.Parameter(ppc.SE3(torch.zeros(4, 7)))
    m = torch.nn.Module()
    m.nodes = p
    assert isinstance(p, torch.nn.Parameter) and p.ltype is ppc.SE3_type and p.requires_grad
    assert list(dict(m.named_parameters())) == ['nodes'] and m.nodes[1:].ltype is ppc.SE3_type


def test_constructors_and_from_matrix():
    assert ppc.SE3([0, 0, 0, 0, 0, 0, 1]).dtype == torch.float32             # train.py:193 style
    assert ppc.SE3(np.zeros((2, 7))).shape == (2, 7)
    with pytest.raises(ppc.IslamError):
        ppc.SE3(torch.zeros(2, 6))
    q = ppc.identity_SO3()
    assert q.ltype is ppc.SO3_type and q.tolist() == [0, 0, 0, 1]
    T = [[0., 1, 0, 0], [0, 0, 1, 0], [1, 0, 0, 0], [0, 0, 0, 1]]            # transformation.py:88-92
    X = ppc.from_matrix(T, ltype=ppc.SE3_type)
    assert torch.allclose(X.matrix(), torch.tensor(T), atol=1e-6)


def test_arithmetic_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    x = ppc.SE3(torch.tensor([[0., 0, 0, 0, 0, 0, 1]]))
    with pytest.raises(ppc.IslamError):
        x.Inv()
    with pytest.raises(ppc.IslamError):
        x @ x


def test_lm_refuses_generic_models():
    with pytest.raises(NotImplementedError):
        ppc.optim.LM(torch.nn.Linear(2, 2))


def test_stop_on_plateau_logic():
    class Opt:
        reject, reject_count, loss, last = 16, 0, 1.0, 2.0
    o = Opt()
    s = ppc.optim.scheduler.StopOnPlateau(o, steps=10, patience=3, decreasing=1e-3)
    for k in range(3):
        o.last, o.loss = o.loss, o.loss - 1e-4
        assert s.continual()
        s.step(o.loss)
    assert not s.continual()                       # three plateau steps in a row
    o2 = Opt()
    s2 = ppc.optim.scheduler.StopOnPlateau(o2, steps=2)
    o2.last, o2.loss = 2.0, 1.0; s2.step(1.0); assert s2.continual()
    o2.last, o2.loss = 1.0, 0.5; s2.step(0.5); assert not s2.continual()
    o3 = Opt(); o3.reject_count = 16
    s3 = ppc.optim.scheduler.StopOnPlateau(o3, steps=10)
    s3.step(1.0)
    assert not s3.continual()
