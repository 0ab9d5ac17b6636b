# DRAM traffic / L2 hit rate of one k_maxsim_tc launch per library build: tools/run_dram.sh "libA.so libB.so" WL
LIBS=$1; WL=${2:-C}
mkdir -p gpurun_out
for lib in $LIBS; do
  COLBERT_B200_LIB=$PWD/$lib timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:k_maxsim_tc -s 1 -c 1 --csv --log-file gpurun_out/dram_$(basename $lib .so)_$WL.csv python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  echo "$lib $WL"; grep -o '"dram__bytes_read.sum[^"]*","[^"]*","[^"]*"\|"lts__t_sector_hit_rate.pct[^"]*","[^"]*","[^"]*"\|"gpu__time_duration.sum[^"]*","[^"]*","[^"]*"' gpurun_out/dram_$(basename $lib .so)_$WL.csv
done
