# ncu --set full of the scoring kernel: tools/run_prof.sh WL [lib.so] [tag]
set -x
mkdir -p gpurun_out
WL=${1:-C}; LIB=${2:-}; TAG=${3:-cur}
[ -n "$LIB" ] && export COLBERT_B200_LIB=$PWD/$LIB
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_maxsim_tc -s 1 -c 1 -o gpurun_out/prof_maxsim_${WL}_$TAG -f python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_${WL}_$TAG.log 2>&1
ls -la gpurun_out | tail -5
