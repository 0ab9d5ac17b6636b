# A/B with options: tools/run_ab2.sh "lib[:key=val[,key=val]] ..." [workloads]   (one rep; prints one line per run)
SPECS=$1; shift
mkdir -p gpurun_out
for spec in $SPECS; do
  lib=${spec%%:*}; opts=""
  if [[ "$spec" == *:* ]]; then for kv in $(echo "${spec#*:}" | tr ',' ' '); do opts="$opts --opt $kv"; done; fi
  for wl in ${@:-C B}; do
    COLBERT_B200_LIB=$PWD/$lib timeout ${BENCH_TMO:-120} python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline --no-extra --no-gate $BENCH_EXTRA $opts > gpurun_out/ab.json 2> gpurun_out/ab.err || { tail -3 gpurun_out/ab.err; head -c 600 gpurun_out/ab.json; }
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab.json"))
    print("$spec $wl", round(d["value"]), "QPS stage34 %.1f ms" % d["roofline"]["stage_ms"]["ms_stage34"], "stage1 %.2f" % d["roofline"]["stage_ms"]["ms_stage1"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "clk/group %.0f smem-pipe %.2f" % ((d["roofline"]["smem_data_pipe"]["clocks_per_group_per_sm"] or 0), (d["roofline"]["smem_data_pipe"]["frac"] or 0)), flush=True)
except Exception as e:
    print("$spec $wl FAILED", e)
PY
  done
done
