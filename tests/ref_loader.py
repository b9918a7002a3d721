"""Test infrastructure: import the reference's own modules (tests/golden/ref_src/, staged by tests/golden/vendor_reference.py)
unmodified on top of `islam_b200.pypose_compat` installed as `pypose`."""
import contextlib
import hashlib
import importlib
import os
import sys
import types

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.path.join(HERE, 'golden', 'ref_src')


@contextlib.contextmanager
def reference_modules():
    """Yields (pvgo, imu_integrator, Datasets.transformation, dense_ba) of the reference, imported over the shim."""
    if not os.path.exists(os.path.join(REF_SRC, 'pvgo.py')):
        pytest.skip('tests/golden/ref_src/ is not staged: run tests/golden/vendor_reference.py (or __graft_entry__.build()) '
                    'in the build container, where /root/reference exists')
    # the files executed are the files the vendoring script saw
    want = dict(l.split()[::-1] for l in open(os.path.join(HERE, 'golden', 'ref_src.sha256')).read().splitlines())
    for rel, digest in want.items():
        assert hashlib.sha256(open(os.path.join(REF_SRC, rel), 'rb').read()).hexdigest() == digest, rel
    import islam_b200.pypose_compat as ppc
    ppc.install()
    # imu_integrator.py:7 imports the CNN-GRU denoiser (a front-end network, out of scope); never instantiated here
    net = types.ModuleType('Network')
    den = types.ModuleType('Network.IMUDenoiseNet')
    den.IMUCorrector_CNN_GRU_WO_COV = type('IMUCorrector_CNN_GRU_WO_COV', (), {})
    net.IMUDenoiseNet = den
    names = ('pvgo', 'imu_integrator', 'Datasets', 'Datasets.transformation', 'dense_ba')
    saved = {k: sys.modules.get(k) for k in ('Network', 'Network.IMUDenoiseNet') + names}
    sys.modules['Network'], sys.modules['Network.IMUDenoiseNet'] = net, den
    for k in names:
        sys.modules.pop(k, None)
    sys.path.insert(0, REF_SRC)
    try:
        mods = (importlib.import_module('pvgo'), importlib.import_module('imu_integrator'),
                importlib.import_module('Datasets.transformation'), importlib.import_module('dense_ba'))
        assert os.path.dirname(mods[0].__file__) == REF_SRC
        yield mods
    finally:
        sys.path.remove(REF_SRC)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
