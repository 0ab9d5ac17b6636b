# A/B builds of the library on the same box: tools/run_ab.sh "libA.so libB.so ..." [workloads]
LIBS=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
for lib in $LIBS; do
for wl in ${@:-C B}; do
  COLBERT_B200_LIB=$PWD/$lib timeout 300 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err || { tail -3 gpurun_out/ab.err; head -c 600 gpurun_out/ab.json; }
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab.json"))
    print("$lib $wl", round(d["value"]), "QPS stage34 %.1f ms" % d["roofline"]["stage_ms"]["ms_stage34"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$lib $wl FAILED", e)
PY
done; done; done
