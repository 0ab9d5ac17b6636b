# tools/run_skeleton.sh: sweep of the pipeline-skeleton microbenchmark (tools/pipe_skeleton.cu); a configuration is flags,accumulators,stages
mkdir -p gpurun_out
O=gpurun_out/${SK_OUT:-r02_pipe_skeleton.txt}
CFGS=${SK_CFGS:-0,2,3 1,2,3 3,2,3 7,2,3 16,2,3 17,2,3 19,2,3 23,2,3 23,4,4}
for cfg in $CFGS; do
  timeout 120 tools/pipe_skeleton ${cfg//,/ } ${SK_N:-80} ${SK_GROUPS:-200000} | tee -a $O
done
