"""Build recipe for libislam_pvgo.so (hand-written CUDA for sm_100a + the C ABI of include/islam_pvgo.h).

    python -m islam_b200.build            # rebuild if any source is newer than the library

nvcc cross-compiles without a GPU; the library is built in-tree (islam_b200/lib/) so it travels to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'lib', 'libislam_pvgo.so')
SOURCES = ['pvgo.cu', 'imu.cu', 'lieops.cu', 'scale.cu', 'symbolic.cpp', 'symbolic3.cpp']
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']


def _nvcc():
    for c in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if c and os.path.exists(c):
            return c
    raise RuntimeError('nvcc not found: libislam_pvgo.so cannot be built')


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'islam_pvgo.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, phase_clocks=False, defines=(), name=None):
    """phase_clocks: developer build (lib/libislam_dbg.so) whose factor kernel stamps clock64() per phase (tools/phase_clocks.py).
    defines / name: further developer builds (extra -D flags, written to lib/<name>)."""
    out = LIB.replace('libislam_pvgo.so', name or 'libislam_dbg.so') if phase_clocks else LIB
    if not force and not phase_clocks and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [_nvcc(), '-O3', '-std=c++17', *ARCH, '-lineinfo', '-Xcompiler', '-fPIC', '-shared',
           '-diag-suppress', '177', '-o', out] + (['-DISLAM_PHASE_CLOCKS'] if phase_clocks else []) + [f'-D{d}' for d in defines] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, '-Xptxas'); cmd.insert(2, '-v')
        print(' '.join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return out


if __name__ == '__main__':
    if '--chain-only' in sys.argv:      # developer experiment: the pivot chain of front4.cuh alone (panel / Schur warps exit)
        print(build(phase_clocks=True, defines=('ISLAM_CHAIN_ONLY',), name='libislam_dbg2.so'))
    else:
        print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, phase_clocks='--phase-clocks' in sys.argv))
