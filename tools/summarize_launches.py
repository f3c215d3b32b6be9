"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (per-kernel share of the step)."""
import collections, csv, re, sys
p = sys.argv[1]
nsteps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
with open(p) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print('kernel | launches/step | us/step | share | avg us')
print('---|---|---|---|---')
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:24]:
    print('%s | %.1f | %.1f | %.1f%% | %.1f' % (k[:80], n / nsteps, t / nsteps, 100 * t / tot, t / n))
print('total | | %.1f us/step (cold-cache, serialised under ncu) | |' % (tot / nsteps))
