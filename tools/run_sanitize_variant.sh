# tools/run_sanitize_variant.sh VARIANT [tool]: compute-sanitizer over the smoke script with an A/B build of the library
mkdir -p gpurun_out
COLBERT_B200_LIB=$PWD/colbert.jl_b200/lib_ab/libcolbert_b200_$1.so timeout ${TMO:-240} compute-sanitizer --tool ${2:-memcheck} --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/sanitize_$1.txt 2>&1
echo "exit $?"; grep -v "^=========     at\|^=========     by\|^=========         in" gpurun_out/sanitize_$1.txt | head -60
