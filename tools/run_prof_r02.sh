# round-2 profile pass at workload C on one GPU: launch list of one step + ncu --set full of the three kernels that matter
set -x
mkdir -p gpurun_out
B="python bench.py --workload ${1:-C} --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-gate"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|^cub|Device' -c 200 --csv --log-file gpurun_out/r02_launches_${1:-C}.csv $B > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_${1:-C}.csv | head -40
for k in k_stage1_tc k_maxsim_tc k_rescore_pairs k_topk_select k_stage2_mark k_bitmap_transpose; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r02_prof_${k}_${1:-C} -f $B > gpurun_out/r02_ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -8
