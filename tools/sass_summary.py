"""Per-kernel SASS mnemonic counts of the shipped library (cuobjdump -sass), the evidence for which hardware paths each
kernel uses: DFMA / DMMA (fp64 pipe / fp64 tensor cores), UBLKCP (TMA bulk copy), LDGSTS (cp.async), SYNCS (mbarrier),
BAR, SHFL, MUFU, LDS/STS, LDG/STG, local-memory spills (LDL/STL).

    python tools/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'islam_b200', 'lib', 'libislam_pvgo.so')
KEYS = ['DFMA', 'DMUL', 'DADD', 'DMMA', 'MUFU', 'SHFL', 'BAR', 'SYNCS', 'UBLKCP', 'LDGSTS', 'UTMALDG', 'UTCHMMA', 'LDS', 'STS', 'LDG', 'STG',
        'LDL', 'STL', 'ATOM', 'RED', 'CCTL', 'ACQBULK', 'UCGABAR']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    arch = set()
    cur = None
    for line in out.splitlines():
        m = re.match(r'\s*arch = (\S+)', line)
        if m:
            arch.add(m.group(1))
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and cur:
            op = m.group(1)
            counts[cur]['_total'] += 1
            base = op.split('.')[0]
            for k in KEYS:
                if base == k or (k in ('LDS', 'STS', 'LDG', 'STG', 'LDL', 'STL') and base == k) or (k == 'ATOM' and base.startswith('ATOM')):
                    counts[cur][k] += 1
    names = list(counts)
    try:
        dm = subprocess.run(['cu++filt'] + names, capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    except Exception:
        pass
    print(f'# {os.path.relpath(LIB, ROOT)}  cubin architectures: {sorted(arch)}')
    print('# kernel | total instructions | ' + ' '.join(KEYS))
    for n, c in counts.items():
        short = demangle.get(n, n).replace('islam::', '').replace('(anonymous namespace)::', '').replace('<unnamed>::', '').replace('(int)', '').replace('(bool)', '')
        short = short.split('(')[0].replace('void ', '')
        print(f'{short:60s} {c["_total"]:7d}  ' + ' '.join(f'{k}={c[k]}' for k in KEYS if c[k]))


if __name__ == '__main__':
    main()
