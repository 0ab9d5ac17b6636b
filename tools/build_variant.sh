#!/bin/bash
# tools/build_variant.sh NAME "-DTC_X=.. ..." : A/B build of the scoring kernel only -> colbert.jl_b200/lib_ab/libcolbert_b200_NAME.so
set -euo pipefail
NAME=$1; DEFS=${2:-}
cd "$(dirname "$0")/../colbert.jl_b200/csrc"
mkdir -p ../lib_ab ../_build_ab
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
$NVCC $FLAGS $DEFS ${PTXAS_V:+-Xptxas -v} -c stage34_tc.cu -o ../_build_ab/stage34_tc_$NAME.o
objs=$(ls ../_build/*.o | grep -v stage34_tc.o)
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../lib_ab/libcolbert_b200_$NAME.so $objs ../_build_ab/stage34_tc_$NAME.o
echo "built lib_ab/libcolbert_b200_$NAME.so"
