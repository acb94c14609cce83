# ncu evidence of round 2 (run under gpurun; every profiled command first runs plain in the same && chain).
# The .ncu-rep files are turned into CSV on the box and deleted: gpurun brings back at most 64 MiB.
set -x
cap() {   # cap <name> <kernel regex> <skip> <count> <source page: 0|1> <command...>
  name=$1; regex=$2; skip=$3; count=$4; src=$5; shift 5
  "$@" > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c $count -o gpurun_out/ncu_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i gpurun_out/ncu_$name.ncu-rep --page raw --csv > gpurun_out/ncu_${name}_raw.csv 2>/dev/null
  if [ "$src" = 1 ]; then ncu -i gpurun_out/ncu_$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/ncu_${name}_source.csv.gz; fi
  rm -f gpurun_out/ncu_$name.ncu-rep
}
B="python bench.py --headline-only --no-cpu-baseline --steps 1 --warmup 3"
cap c4_f64 sweep_group_kernel 4 1 1 $B
cap c4_f32 sweep_tc_kernel 4 1 1 $B --dtype f32
cap c1_small small_n_kernel 3 1 0 python tools/small_n_bench.py
cap c3_f64 sweep_kernel 3 1 0 python tools/small_n_bench.py
cap aux "gemm_nt|chol_block|probe_kernel|grad_kernel" 0 12 0 python tools/profile_aux.py
cap wk_c3 sweep_warp_kernel 3 1 1 env BOPY_B200_WARP_KERNEL=1 python tools/small_n_bench.py
L="python bench.py --steps 2 --warmup 3 --c5-candidates 131072 --cpu-budget 1"
$L > gpurun_out/plain_launches.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/ncu_launches_r02_bench.csv $L > gpurun_out/ncu_launches.log 2>&1
du -sh gpurun_out
echo done
