#!/bin/bash
# Developer build of a VARIANT of the library for same-box A/B runs: tools/build_variant_lib.sh NAME "-DFLAG=.." [file.cu ...]
# recompiles the listed sources (default: attn_tc.cu) with the extra flags, links them with the product's other objects
# (pianobart_b200/build/*.o - run `python -c "import __graft_entry__ as g; g.build()"` first) into
# pianobart_b200/libpianobart_b200_NAME.so; select it with PIANOBART_B200_LIB=...  The product library is never built this way.
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2; shift 2
srcs=${@:-attn_tc.cu}
mkdir -p pianobart_b200/build_trace/$name
objs=""
for o in pianobart_b200/build/*.o; do
  b=$(basename ${o%.o})
  skip=0
  for f in $srcs; do [ "${f%.cu}" = "$b" ] && skip=1; done
  [ $skip = 0 ] && objs="$objs $o"
done
for f in $srcs; do
  o=pianobart_b200/build_trace/$name/${f%.cu}.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c pianobart_b200/csrc/$f -o $o &
  objs="$objs $o"
done
wait
/usr/local/cuda/bin/nvcc -shared -o pianobart_b200/libpianobart_b200_$name.so $objs -gencode arch=compute_100a,code=sm_100a
echo built pianobart_b200/libpianobart_b200_$name.so
