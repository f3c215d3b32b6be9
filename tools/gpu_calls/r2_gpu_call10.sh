mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-decode --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_launches.csv')) if len(r)>10 and r[0].isdigit()]
print(len(rows),'launches captured')
PY
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 700 -c 12 -o gpurun_out/r2_gemm python bench.py --steps 1 --warmup 1 --no-decode --no-cpu-baseline > gpurun_out/r2_ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"octuple_front_fwd|octuple_onehot|attn_fwd_kernel|attn_bwd_dkv|attn_bwd_dq" -s 20 -c 8 -o gpurun_out/r2_front_attn python bench.py --steps 1 --warmup 1 --no-decode --no-cpu-baseline > gpurun_out/r2_ncu_front.log 2>&1
ls -la gpurun_out/*.ncu-rep
