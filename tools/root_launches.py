"""Summary of an ncu launch list (--metrics gpu__time_duration.sum --csv) restricted to the dense-root kernels."""
import csv, collections, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
tot = collections.defaultdict(float); cnt = collections.Counter(); mx = collections.defaultdict(float)
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    k = row['Kernel Name'].split('(')[0]
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e6 if u in ('ns', 'nsecond') else (v / 1e3 if u in ('us', 'usecond') else v)
    tot[k] += v; cnt[k] += 1; mx[k] = max(mx[k], v)
for k in sorted(tot, key=tot.get, reverse=True):
    print(f'{k:24s} n={cnt[k]:5d} total {tot[k]:9.2f} ms  avg {1e3 * tot[k] / cnt[k]:8.1f} us  max {1e3 * mx[k]:8.1f} us')
