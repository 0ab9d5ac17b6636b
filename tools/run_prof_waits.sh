# tools/run_prof_waits.sh "variant ..." : wait/busy split per role of -DTC_PROF=1 builds at workload C
mkdir -p gpurun_out
for v in $1; do
  CB_TC_PROF_OUT=$PWD/gpurun_out/tcprof_$v.npy COLBERT_B200_LIB=$PWD/colbert.jl_b200/lib_ab/libcolbert_b200_$v.so timeout ${BENCH_TMO:-120} python bench.py --workload ${WL:-C} --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-gate ${BENCH_EXTRA:-} > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  echo "== $v  $(python -c "import json;d=json.load(open('gpurun_out/ab.json'));print('stage34 %.1f ms'%d['roofline']['stage_ms']['ms_stage34'], d['clocks']['sm_mhz'])")" | tee -a gpurun_out/r02_wait_profile.txt
  python tools/tc_wait_profile.py gpurun_out/tcprof_$v.npy ${GROUPS_PER_STEP:-40400000} | tee -a gpurun_out/r02_wait_profile.txt
done
