# tools/run_ab3.sh: A/B of scoring-kernel variants at workload C (+ parity subset for the new ones)
mkdir -p gpurun_out
L=colbert.jl_b200/lib_ab
for v in ${PARITY_VARIANTS:-}; do
  echo "== parity $v"; COLBERT_B200_LIB=$PWD/$L/libcolbert_b200_$v.so timeout ${PYTEST_TMO:-300} python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
done
SPECS=""; for v in $VARIANTS; do SPECS="$SPECS $L/libcolbert_b200_$v.so"; done
bash tools/run_ab2.sh "$SPECS" ${WORKLOADS:-C} 2>&1 | tee -a gpurun_out/${OUT:-r02_ab3.txt}
