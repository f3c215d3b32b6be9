mkdir -p gpurun_out
# launch list of the batch-64 decode step (cold-cache, serialised: shares only)
PIANOBART_B200_DECODE_STEPS=4 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_decode_b64_launches.csv python tools/gpu_decode_bench.py 64 > gpurun_out/r2_decode_b64_ncu.log 2>&1
tail -3 gpurun_out/r2_decode_b64_ncu.log
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_decode_b64_launches.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4][:60]; v=float(r[-1].replace(',',''))
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]:
    print('%-62s n=%5d  total %10.1f us  avg %8.2f us  %5.1f%%'%(k,a[0],a[1]/1e3,a[1]/a[0]/1e3,100*a[1]/tot))
PY
