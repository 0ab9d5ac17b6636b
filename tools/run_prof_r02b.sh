# launch list of one step (our kernels only) + ncu --set full of the TMEM-A build's scoring kernel
mkdir -p gpurun_out
B="python bench.py --workload C --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-gate"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 120 --csv --log-file gpurun_out/r02_launches_C.csv $B > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_C.csv | head -40
COLBERT_B200_LIB=$PWD/colbert.jl_b200/lib_ab/libcolbert_b200_tmemA.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_maxsim_tc -s 1 -c 1 -o gpurun_out/r02_prof_k_maxsim_tc_tmemA_C -f $B > gpurun_out/r02_ncu_tmemA.log 2>&1
ls -la gpurun_out/r02_prof_k_maxsim_tc_tmemA_C.ncu-rep
