# compute-sanitizer passes over tools/sanitize_smoke.py (tiny index, every kernel family).  Output -> gpurun_out/r02_sanitizer_<tool>.txt
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/r02_sanitizer_$tool.txt 2>&1
  echo "== $tool exit $?"; tail -6 gpurun_out/r02_sanitizer_$tool.txt
done
