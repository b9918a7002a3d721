"""pypose.optim.corrector — imported (unused) by /root/reference/pvgo.py:10."""


class FastTriggs:
    def __init__(self, *a, **kw):
        raise NotImplementedError('correctors are not used by iSLAM (pvgo.py:171 passes none)')


Triggs = FastTriggs
