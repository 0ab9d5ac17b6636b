# ncu --set full of the shipped k_maxsim_tc at workloads C and B (final build of the round)
mkdir -p gpurun_out
for wl in C B; do
  B="python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-gate"
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_maxsim_tc -s 1 -c 1 -o gpurun_out/r02_prof_k_maxsim_tc_${wl}_v3 -f $B > gpurun_out/r02_ncu_k_maxsim_tc_${wl}_v3.log 2>&1
  ls -la gpurun_out/r02_prof_k_maxsim_tc_${wl}_v3.ncu-rep
done
