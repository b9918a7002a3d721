"""The C-ABI library loads on a machine without a GPU and exports every symbol include/islam_pvgo.h declares."""
import ctypes as C
import os
import re

import pytest

from islam_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, 'include', 'islam_pvgo.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(islam_[a-z0-9_]+)\s*\(', src)))


def test_library_is_built_in_tree():
    assert os.path.exists(build.LIB) or build.build()
    assert os.path.exists(_lib.LIB_PATH)


def test_every_declared_symbol_is_exported_and_bound():
    names = _header_functions()
    assert len(names) >= 35
    L = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f'{n} is declared in include/islam_pvgo.h but not exported'
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)
    assert b'sm_100a' in _lib.lib().islam_version()


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.PvgoOpts) == 32
    assert C.sizeof(_lib.LMState) == 10 * 8 + 16 * 4
    assert C.sizeof(_lib.LMParams) == 10 * 8 + 4 * 4 + 8
    assert C.sizeof(_lib.PvgoDims) == 12 * 4 + 3 * 8 + 8


def test_product_fails_loudly_without_gpu_or_library(monkeypatch):
    import torch
    from islam_b200.pvgo import run_pvgo
    from islam_b200.solver import PVGOSolver
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            PVGOSolver(4, [[0, 1], [1, 2], [2, 3]], device='cpu')
        with pytest.raises(_lib.IslamError):
            run_pvgo(torch.zeros(3, 7), torch.zeros(3, 3), torch.zeros(2, 7), torch.tensor([[0, 1], [1, 2]]),
                     torch.zeros(2), torch.zeros(2, 4), torch.zeros(2, 3), torch.zeros(2, 3))
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libislam_pvgo.so')
    with pytest.raises(_lib.IslamError):
        _lib.lib()


def test_sass_is_sm100a():
    """The shipped cubin targets sm_100a only (no PTX JIT path for another architecture)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    out = subprocess.run([cuobjdump, '-lelf', _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out and 'sm_90' not in out
