set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_pytest_a.log
tail -15 gpurun_out/r2_pytest_a.log
( PIANOBART_B200_LIB=$PWD/pianobart_b200/libpianobart_b200_single_phase_demo.so timeout 300 python tools/attn_late_tile_demo.py; echo "exit $?"; timeout 300 python tools/attn_late_tile_demo.py; echo "exit $?" ) > gpurun_out/r2_late_tile_demo.log 2>&1
cat gpurun_out/r2_late_tile_demo.log | grep -v "^frame\|^  File" | head -40
( MAXFAIL=2 bash tools/cold_start_diag.sh pianobart_b200/libpianobart_b200_r1.so 36 r1 coredump
  PB_IDLE_GAPS=60 STEPS=2 bash tools/cold_start_diag.sh pianobart_b200/libpianobart_b200_r1.so 2 r1gaps coredump
  bash tools/cold_start_diag.sh default 60 r2
  PB_IDLE_GAPS=60 STEPS=2 bash tools/cold_start_diag.sh default 2 r2gaps ) > gpurun_out/r2_cold_start.log 2>&1
cat gpurun_out/r2_cold_start.log
ls -la gpurun_out | grep core
