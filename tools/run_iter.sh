# quick iteration: parity tests of the scoring path + short benches (no CPU baseline leg)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for wl in C B; do
  timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline $BENCH_EXTRA > gpurun_out/iter_$wl.json 2> gpurun_out/iter_$wl.err
  python - <<PY
import json
d=json.load(open("gpurun_out/iter_$wl.json"))
print("$wl", round(d["value"]), "QPS", d["roofline"]["stage_ms"], "hbm_equiv", round(d["roofline"]["hbm_equiv"]["frac"],3), d["clocks"])
PY
done
