#!/bin/bash
# Builds libcolbert_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
OUT=../lib
mkdir -p "$OUT" ../_build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function --expt-relaxed-constexpr ${CB_EXTRA_NVCC_FLAGS:-}"
pids=()
for f in index index_open multi stage1 stage1_tc stage2 stage34_generic stage34_tc stage5 search hooks plaid; do
  if [ ! -f ../_build/$f.o ] || [ $f.cu -nt ../_build/$f.o ] || [ ptx.cuh -nt ../_build/$f.o ] || [ common.cuh -nt ../_build/$f.o ] || [ ../../include/colbert_b200.h -nt ../_build/$f.o ] || [ jld2.h -nt ../_build/$f.o ]; then
    $NVCC $FLAGS -c $f.cu -o ../_build/$f.o &
    pids+=($!)
  fi
done
if [ ! -f ../_build/jld2.o ] || [ jld2.cpp -nt ../_build/jld2.o ] || [ jld2.h -nt ../_build/jld2.o ]; then
  g++ -O2 -std=c++17 -fPIC -Wall -c jld2.cpp -o ../_build/jld2.o &
  pids+=($!)
fi
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libcolbert_b200.so" ../_build/*.o
echo "built $OUT/libcolbert_b200.so"
