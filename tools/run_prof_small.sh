# ncu --set full of the stage 1 / 2 / 5 kernels (one launch each), workload C
set -x
mkdir -p gpurun_out
for kn in k_stage1_tc k_stage2_mark k_topk_select k_stage1_rescore; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s 1 -c 1 -f -o gpurun_out/prof_${kn}_C python bench.py --workload C --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_${kn}.log 2>&1
done
ls -la gpurun_out | grep prof_k_
