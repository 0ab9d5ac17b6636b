# PLAID leg (BASELINE config 5) at workload C: launch list of one step (our kernels only) + ncu --set full of k_plaid_approx
mkdir -p gpurun_out
B="python tools/bench_plaid.py --workload C --steps 2 --warmup 1 --parity-queries 0 --skip-exhaustive"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -s 60 -c 160 --csv --log-file gpurun_out/r02_launches_plaid_C.csv $B > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_plaid_C.csv | head -40 | tee gpurun_out/r02_launch_summary_plaid_C.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_plaid_approx -s 1 -c 1 -o gpurun_out/r02_prof_k_plaid_approx_C -f $B > gpurun_out/r02_ncu_k_plaid_approx.log 2>&1
ls -la gpurun_out/r02_prof_k_plaid_approx_C.ncu-rep
