"""pypose.optim.strategy — TrustRegion parameters (applied on the device, csrc/lm.cuh lm_control)."""


class TrustRegion:
    """strategy.TrustRegion(radius=1e6, high=.5, low=1e-3, up=2., down=.5, factor=.5, min=1e-6, max=1e16): the constants
    travel to the device through islam_lm_params (include/islam_pvgo.h)."""

    def __init__(self, radius=1e6, high=.5, low=1e-3, up=2., down=.5, factor=.5, min=1e-6, max=1e16):
        assert (high > low) and (up > 1) and (down < 1), 'invalid trust-region constants'
        self.radius, self.high, self.low, self.up, self.down = radius, high, low, up, down
        self.factor, self.min, self.max = factor, min, max

    def lm_params(self):
        return dict(radius=float(self.radius), high=float(self.high), low=float(self.low), up=float(self.up),
                    down=float(self.down), factor=float(self.factor), tr_min=float(self.min), tr_max=float(self.max))


class Constant:
    def __init__(self, *a, **kw):
        raise NotImplementedError('only strategy.TrustRegion (pvgo.py:170) is provided on the B200 path')


Adaptive = Constant
