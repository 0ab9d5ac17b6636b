# launch list of the PLAID leg at workload C (our kernels only)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -s 60 -c 160 --csv --log-file gpurun_out/r02_launches_plaid_C.csv python tools/bench_plaid.py --workload C --steps 2 --warmup 1 --parity-queries 0 --skip-exhaustive > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_plaid_C.csv | head -40
