"""pypose.optim.kernel — imported (unused) by /root/reference/pvgo.py:9; only the trivial kernel exists here."""


class Trivial:
    pass


class Huber:
    def __init__(self, *a, **kw):
        raise NotImplementedError('robust kernels are not used by iSLAM (pvgo.py:171 passes none)')


PseudoHuber = Cauchy = Huber
