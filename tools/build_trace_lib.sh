#!/bin/bash
# Developer build of the library with the clock64 phase traces compiled in (-DPB_TRACE) -> pianobart_b200/libpianobart_b200_trace.so
# (use with PIANOBART_B200_LIB=pianobart_b200/libpianobart_b200_trace.so; the product library is never built with it)
set -e
cd "$(dirname "$0")/.."
mkdir -p pianobart_b200/build_trace
objs=""
for f in pianobart_b200/csrc/*.cu; do
  o=pianobart_b200/build_trace/$(basename ${f%.cu}).o
  if [ ! -f $o ] || [ $f -nt $o ] || [ pianobart_b200/csrc/ptx.cuh -nt $o ]; then
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DPB_TRACE $PB_TRACE_EXTRA -c $f -o $o &
  fi
  objs="$objs $o"
done
wait
/usr/local/cuda/bin/nvcc -shared -o pianobart_b200/libpianobart_b200_trace.so $objs -gencode arch=compute_100a,code=sm_100a
echo built pianobart_b200/libpianobart_b200_trace.so
