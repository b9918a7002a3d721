"""pypose.optim.strategy — TrustRegion parameters (applied on the device, csrc/lm.cuh k_lm_control)."""


class TrustRegion:
    def __init__(self, radius=1e6, high=.5, low=1e-3, up=2., down=.5, factor=.5, min=1e-6, max=1e16):
        if (high, low, up, down, factor, min, max) != (.5, 1e-3, 2., .5, .5, 1e-6, 1e16):
            raise NotImplementedError('non-default TrustRegion constants are not wired through the shim')
        self.radius = radius


class Constant:
    def __init__(self, *a, **kw):
        raise NotImplementedError('only strategy.TrustRegion (pvgo.py:170) is provided on the B200 path')


Adaptive = Constant
