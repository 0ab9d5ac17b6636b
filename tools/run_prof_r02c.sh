# final round-2 profile pass of the shipped build at workload C: launch list of one step (our kernels only) + ncu --set full of k_maxsim_tc
mkdir -p gpurun_out
B="python bench.py --workload C --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-gate"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 120 --csv --log-file gpurun_out/r02_launches_C_v2.csv $B > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_C_v2.csv | head -40 | tee gpurun_out/r02_launch_summary_C_v2.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_maxsim_tc -s 1 -c 1 -o gpurun_out/r02_prof_k_maxsim_tc_C_v2 -f $B > gpurun_out/r02_ncu_k_maxsim_tc_v2.log 2>&1
ls -la gpurun_out/r02_prof_k_maxsim_tc_C_v2.ncu-rep
