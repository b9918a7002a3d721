"""pypose.optim.solver — configuration tokens; the linear algebra itself is the multifrontal Cholesky in csrc/solver.cuh."""


class Cholesky:
    def __init__(self, upper=False):
        self.upper = upper


class PINV:
    def __init__(self, *a, **kw):
        raise NotImplementedError('only solver.Cholesky (pvgo.py:169) is provided on the B200 path')


LSTSQ = PINV
