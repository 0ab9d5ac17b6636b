mkdir -p gpurun_out
v=tr_iss
CB_TC_PROF_OUT=$PWD/gpurun_out/tcprof_$v.npy COLBERT_B200_LIB=$PWD/colbert.jl_b200/lib_ab/libcolbert_b200_$v.so timeout 120 python bench.py --workload C --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-gate > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python - <<'PY'
import numpy as np, json
d=json.load(open("gpurun_out/ab.json")); print("stage34 %.1f ms"%d["roofline"]["stage_ms"]["ms_stage34"], d["clocks"]["sm_mhz"])
t=np.load("gpurun_out/tcprof_tr_iss_trace.npy").astype(np.int64)
ok=(t[0]>0)&(t[1]>0)&(t[2]>0); t=t[:,ok]; print(ok.sum(),"groups")
def st(n,v): print(f"{n:40s} median {np.median(v):7.0f} mean {v.mean():7.0f} p10 {np.percentile(v,10):7.0f} p90 {np.percentile(v,90):7.0f}")
st("period a_full(g)->a_full(g+1)", np.diff(t[0]))
st("a_full granted -> d_empty granted", t[1]-t[0])
st("d_empty granted -> issued (8 MMA+commits)", t[2]-t[1])
st("issued(g) -> a_full granted (g+1)", t[0][1:]-t[2][:-1])
b=t[0][100]
for g in range(100,120): print(g, t[0][g]-b, t[1][g]-b, t[2][g]-b)
PY
