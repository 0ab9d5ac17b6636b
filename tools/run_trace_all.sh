mkdir -p gpurun_out
v=${1:-tr_all}
CB_TC_PROF_OUT=$PWD/gpurun_out/tcprof_$v.npy COLBERT_B200_LIB=$PWD/colbert.jl_b200/lib_ab/libcolbert_b200_$v.so timeout 120 python bench.py --workload C --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-gate ${BENCH_EXTRA:-} > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "import json;d=json.load(open('gpurun_out/ab.json'));print('stage34 %.1f ms'%d['roofline']['stage_ms']['ms_stage34'], d['clocks']['sm_mhz'])"
python tools/tc_timeline.py gpurun_out/tcprof_${v}_trace.npy | tee gpurun_out/r02_timeline_$v.txt
