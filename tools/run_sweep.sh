# sweep a library option on one workload: tools/run_sweep.sh WL key v1 v2 ...
WL=$1; KEY=$2; shift 2
mkdir -p gpurun_out
for v in "$@"; do
  timeout 300 python bench.py --workload $WL --steps 3 --warmup 2 --no-cpu-baseline --opt $KEY=$v > gpurun_out/sweep_${WL}_${KEY}_$v.json 2> gpurun_out/sweep.err || tail -3 gpurun_out/sweep.err
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep_${WL}_${KEY}_$v.json"))
print("$WL $KEY=$v", round(d["value"]), "QPS stage34 %.1f ms" % d["roofline"]["stage_ms"]["ms_stage34"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
