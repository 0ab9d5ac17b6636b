#!/bin/bash
# Builds a variant of the library for A/B runs: build_variant.sh NAME "extra nvcc flags"
# -> ../lib_ab/libcolbert_b200_NAME.so (git-ignored; travels with gpurun).  See tools/run_ab.sh.
set -euo pipefail
cd "$(dirname "$0")"
NAME=$1; EXTRA=${2:-}
mkdir -p ../lib_ab ../_build_ab/$NAME
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $EXTRA"
for f in index index_open multi stage1 stage1_tc stage2 stage34_generic stage34_tc stage5 search hooks plaid; do
  $NVCC $FLAGS -c $f.cu -o ../_build_ab/$NAME/$f.o &
done
g++ -O2 -std=c++17 -fPIC -c jld2.cpp -o ../_build_ab/$NAME/jld2.o
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../lib_ab/libcolbert_b200_$NAME.so ../_build_ab/$NAME/*.o
echo "built ../lib_ab/libcolbert_b200_$NAME.so"
