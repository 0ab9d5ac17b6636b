# Round-end GPU pass: parity tests, smoke, benches (C = headline, B), launch list, ncu full captures.
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --workload C --steps 5 --warmup 3 > gpurun_out/bench_C.json 2> gpurun_out/bench_C.err; tail -c 2500 gpurun_out/bench_C.json
timeout 600 python bench.py --workload B --steps 5 --warmup 3 > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err; tail -c 600 gpurun_out/bench_B.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/launches_C.csv python bench.py --workload C --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_C.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_maxsim_tc -s 1 -c 1 -f -o gpurun_out/prof_maxsim_C_final python bench.py --workload C --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_C.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_maxsim_tc -s 1 -c 1 -f -o gpurun_out/prof_maxsim_B_final python bench.py --workload B --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_B.log 2>&1
ls -la gpurun_out | tail -12
