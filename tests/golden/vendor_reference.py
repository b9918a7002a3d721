"""Test infrastructure: stage the reference's own back-end modules where the GPU box can import them.

    python tests/golden/vendor_reference.py        (also called by __graft_entry__.build())

/root/reference exists only in the build container.  The `-m gpu` test tests/test_gpu_reference_files.py executes
the UNMODIFIED files  pvgo.py, imu_integrator.py, dense_ba.py, Datasets/transformation.py  on top of `islam_b200.pypose_compat`
(installed as `pypose`), so byte-identical copies are placed under tests/golden/ref_src/ — a directory that is
git-ignored (reference sources never enter this repository's history) but not gpurun-ignored, so it travels to the box
with the snapshot exactly like the built .so files.  A sha256 manifest of what was staged is written next to them;
the committed manifest (ref_src.sha256) lets the test assert that it ran the files this script saw.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, 'ref_src')
REF = os.environ.get('ISLAM_REFERENCE', '/root/reference')
FILES = {'pvgo.py': 'pvgo.py', 'imu_integrator.py': 'imu_integrator.py', 'dense_ba.py': 'dense_ba.py',
         os.path.join('Datasets', 'transformation.py'): os.path.join('Datasets', 'transformation.py')}


def sha(path):
    return hashlib.sha256(open(path, 'rb').read()).hexdigest()


def vendor(quiet=False):
    if not os.path.isdir(REF):
        if not quiet:
            print(f'{REF} is not present: nothing staged (the GPU box uses what the build container staged)')
        return False
    lines = []
    for src, dst in FILES.items():
        d = os.path.join(DST, dst)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(REF, src), d)
        lines.append(f'{sha(d)}  {dst}')
    open(os.path.join(DST, 'Datasets', '__init__.py'), 'w').close()
    manifest = '\n'.join(lines) + '\n'
    open(os.path.join(DST, 'MANIFEST.sha256'), 'w').write(manifest)
    committed = os.path.join(HERE, 'ref_src.sha256')
    if not os.path.exists(committed) or open(committed).read() != manifest:
        open(committed, 'w').write(manifest)
    if not quiet:
        print(manifest, end='')
    return True


if __name__ == '__main__':
    sys.exit(0 if vendor() else 1)
