"""Top stall sites of one kernel from `ncu -i X.ncu-rep --page source --csv` output (SASS view).
usage: python tools/ncu_top_sass.py file.csv [N] [kernel index]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sec = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # which kernel of the file
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
rows = rows[starts[sec]:starts[sec + 1]]
h = [i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r][0]
hdr = rows[h]
si, sa, ex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
stalls = [k for k, x in enumerate(hdr) if x.startswith('stall_') and 'Not Issued' not in x]
data = [r for r in rows[h + 1:] if len(r) > sa]
tot = sum(float(r[sa] or 0) for r in data)
print(rows[0][1][:120], 'samples', tot, 'instr', sum(float(r[ex] or 0) for r in data))
idx = {id(r): i for i, r in enumerate(data)}
for r in sorted(data, key=lambda r: -float(r[sa] or 0))[:n]:
    why = sorted(((float(r[k] or 0), hdr[k][6:]) for k in stalls), reverse=True)[:2]
    print('%5d %5.1f%% ex=%-7s %-70s %s' % (idx[id(r)], 100 * float(r[sa]) / tot, r[ex], r[si][:70],
                                          ' '.join('%s:%d' % (w, v) for v, w in why if v > 0)))
