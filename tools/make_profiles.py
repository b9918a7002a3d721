"""Turns the raw artefacts of tools/gpu_final.sh (gpurun_out/) into the committed summaries under profiles/.
Runs on the CPU box (needs ncu only to read the .ncu-rep and nvdisasm for the line table)."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')
TAG = sys.argv[1] if len(sys.argv) > 1 else 'r01'


def copy_json(src, dst):
    p = os.path.join(OUT, src)
    if not os.path.exists(p):
        return
    for line in open(p):
        if line.startswith('{'):
            open(os.path.join(PROF, dst), 'w').write(line)


def launches():
    p = os.path.join(OUT, 'launches.csv')
    if not os.path.exists(p):
        return
    rows = [r for r in csv.reader(open(p)) if len(r) > 10]
    ci = {h: i for i, h in enumerate(rows[0])}
    per = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ci['Kernel Name']].split('(')[0].replace('void ', '').replace('islam::', '')
        val = float(r[ci['Metric Value']].replace(',', ''))
        unit = r[ci['Metric Unit']]
        if r[ci['Metric Name']] == 'gpu__time_duration.sum' and unit == 'us':
            val *= 1e3
        if unit == 'Kbyte': val *= 1e3
        if unit == 'Mbyte': val *= 1e6
        per.setdefault((int(r[ci['ID']]), name, r[ci['Grid Size']]), {})[r[ci['Metric Name']]] = val
    items = list(per.items())
    # one try = from k_begin_try to the launch before the next k_begin_try
    starts = [i for i, (k, _) in enumerate(items) if k[1] == 'k_begin_try']
    if len(starts) >= 2:
        items = items[starts[0]:starts[1]]
    tot = sum(m['gpu__time_duration.sum'] for _, m in items)
    with open(os.path.join(PROF, f'{TAG}_try_launches_c2.csv'), 'w') as f:
        f.write('# ncu launch list of ONE LM try inside `python bench.py --steps 2 --warmup 1` (C2), --clock-control none\n')
        f.write('# per-launch times are cold-cache, serialised and WITHOUT the programmatic-dependent-launch overlap: compare SHARES\n')
        f.write('kernel,grid,time_ns,share,dram_read,dram_write\n')
        agg = collections.OrderedDict()
        for (_, name, grid), m in items:
            f.write(f"{name},{grid.replace(',', ' ')},{int(m['gpu__time_duration.sum'])},{m['gpu__time_duration.sum'] / tot:.3f},"
                    f"{int(m.get('dram__bytes_read.sum', 0))},{int(m.get('dram__bytes_write.sum', 0))}\n")
            a = agg.setdefault(name.split('<')[0], [0, 0.0, 0.0, 0.0])
            a[0] += 1; a[1] += m['gpu__time_duration.sum']; a[2] += m.get('dram__bytes_read.sum', 0); a[3] += m.get('dram__bytes_write.sum', 0)
        f.write('# per kernel family\n')
        for name, a in agg.items():
            f.write(f'# {name},{a[0]} launches,{int(a[1])} ns,{a[1] / tot:.3f},{int(a[2])},{int(a[3])}\n')
        f.write(f'# TOTAL,{len(items)} launches,{int(tot)} ns\n')
    fac = [m for (_, name, _), m in items if name.startswith('k_factor3')]
    bs = [m for (_, name, _), m in items if name.startswith('k_backsolve3')]
    if fac:
        fb = sum(m.get('dram__bytes_read.sum', 0) + m.get('dram__bytes_write.sum', 0) for m in fac)
        bb = sum(m.get('dram__bytes_read.sum', 0) + m.get('dram__bytes_write.sum', 0) for m in bs)
        json.dump({'source': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none (cold cache per replay) on one '
                             'LM try of `python bench.py --steps 2 --warmup 1`, C2',
                   'k_factor_level_bytes_per_factorisation': fb, 'k_factor_level_bytes_per_launch': fb / len(fac),
                   'k_backsolve_level_bytes_per_solve': bb, 'algorithmic_bytes_per_factorisation_survey_8d': 21.6e6},
                  open(os.path.join(PROF, 'traffic.json'), 'w'), indent=1)


def ncu_full():
    rep = os.path.join(OUT, 'f3_l3.ncu-rep')
    if not os.path.exists(rep):
        return
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    keep = re.compile(r'^(Kernel Name|Grid Size|Block Size|gpu__time_duration\.sum|launch__registers_per_thread|launch__occupancy_limit.*|'
                      r'sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__issue_active\.avg\.pct_of_peak_sustained_active|'
                      r'sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|'
                      r'l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum(\.pct_of_peak_sustained_elapsed)?|'
                      r'l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|dram__bytes_(read|write)\.sum|lts__t_bytes\.sum|'
                      r'dram__throughput\.avg\.pct_of_peak_sustained_elapsed|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|'
                      r'smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|smsp__inst_executed\.sum|sm__cycles_elapsed\.max)$')
    with open(os.path.join(PROF, f'{TAG}_factor3_full_summary.csv'), 'w') as f:
        f.write('# ncu --set full --clock-control none of ONE k_factor3<512,1,2> launch (64 interior separator fronts, level 3 of C2)\n')
        f.write('metric,unit,value\n')
        for h, u, v in zip(hdr, units, vals):
            if keep.match(h):
                f.write(f'"{h}","{u}","{v}"\n')
    # stall samples per phase of the kernel (SASS samples joined with nvdisasm's line table)
    lib = os.path.join(ROOT, 'islam_b200', 'lib', 'libislam_pvgo.so')
    tmp = '/tmp/islam_cubin'
    os.makedirs(tmp, exist_ok=True)
    subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=tmp, capture_output=True)
    cubin = os.path.join(tmp, 'pvgo.sm_100a.cubin')
    if not os.path.exists(cubin):
        return
    sass = subprocess.run(['nvdisasm', '--print-line-info', cubin], capture_output=True, text=True).stdout.split('\n')
    start = next((i for i, l in enumerate(sass) if l.startswith('//--------------------- .text._ZN5islam9k_factor3ILi512ELi1ELi2')), None)
    if start is None:
        return
    addr2line, cur = {}, None
    for l in sass[start + 1:]:
        if l.startswith('//--------------------- '):
            break
        m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            addr2line[int(m.group(1), 16)] = cur
    src_csv = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(src_csv.splitlines()))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    fl = lambda x: float(x) if x else 0.0
    base = int(data[0][ci['Address']], 16)
    src = open(os.path.join(ROOT, 'islam_b200', 'csrc', 'solver3.cuh')).read().split('\n')
    find = lambda s: next((i + 1 for i, l in enumerate(src) if s in l), None)
    marks = [('preamble (zero fill, map resolution, pre-dependency A1)', 1), ('grid dependency + late A1 / stage 2', find('cudaGridDependencySynchronize();           // previous level')),
             ('A2 extend-add of the children', find('// A2. extend-add')), ('stage-1 dump', find('if (stage == 1) {                          // dump')),
             ('diagonal-block chain (warp 0)', find('auto diag_block = [&]')), ('store of finished block columns', find('auto store_block = [&]')),
             ('Schur update U -= L21 L21^T', find('// U -= L21[:, blocks]')), ('step loop: Linv store, row solve, barriers', find('if (warp == 0) diag_block(0, false, sLinv);')),
             ('trailing panel update', find('// trailing update: column block cb only')), ('tail (U store)', find('if (!ok && tid == 0) *chol_fail = 1;'))]
    marks = [m for m in marks if m[1]]
    agg, bar, inst = collections.Counter(), collections.Counter(), collections.Counter()
    for r in data:
        ln = addr2line.get(int(r[ci['Address']], 16) - base)
        n = fl(r[ci['# Samples']])
        name = '(inlined intrinsics)' if not ln or ln[0] != 'solver3.cuh' else [m[0] for m in marks if m[1] <= ln[1]][-1]
        agg[name] += n; bar[name] += fl(r[ci['stall_barrier']]); inst[name] += fl(r[ci['Instructions Executed']])
    tot = sum(agg.values())
    with open(os.path.join(PROF, f'{TAG}_factor3_stall_by_phase.txt'), 'w') as f:
        f.write('# warp-state samples of the same ncu capture, grouped by phase of k_factor3 (SASS samples joined with the nvdisasm line table)\n')
        for k, v in agg.items():
            f.write(f'{k:60s} samples {int(v):5d} ({100 * v / max(tot, 1):4.1f} %)  of which barrier-stalled {int(bar[k]):4d}  warp instructions {int(inst[k]):8d}\n')


def copy_text(src, dst, header):
    p = os.path.join(OUT, src)
    if os.path.exists(p):
        open(os.path.join(PROF, dst), 'w').write(header + open(p).read())


if __name__ == '__main__':
    os.makedirs(PROF, exist_ok=True)
    copy_json('bench_n1.json', f'{TAG}_bench_n1.json')
    copy_json('bench_ref.json', f'{TAG}_bench_reference.json')
    launches()
    ncu_full()
    copy_text('timeline.log', f'{TAG}_level_timeline.txt', '# tools/level_timeline.py: globaltimer per front of one C2 factorisation (libislam_dbg.so)\n')
    copy_text('phase.log', f'{TAG}_factor_phase_clocks.txt', '# tools/phase_clocks.py: clock64() stamps of the middle CTA of a level, per grid size (libislam_dbg.so)\n')
    copy_text('scale_bench.log', f'{TAG}_scale_microbench.txt', '# tools/scale_bench.py: islam_scale_from_disp_flow (SURVEY 8f rank 4)\n')
    copy_text('configs.log', f'{TAG}_configs.txt', '# tools/configs_bench.py: every BASELINE.json config on one B200\n')
    print(sorted(os.listdir(PROF)))
