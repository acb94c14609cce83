# ncu capture of sweep_group_kernel at the bench size for one (G:S:lead) setting; the .ncu-rep is turned into CSV and deleted
# (gpurun brings back at most 64 MiB).   bash tools/ncu_group_mode.sh 8:3:8 [C4] [21]
cfg=${1:-default}; key=${2:-C4}; mlog=${3:-21}; name=group_${key}_${cfg//:/_}
python tools/group_mode_bench.py $key $mlog $cfg --once > gpurun_out/plain_$name.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_group_kernel -s 1 -c 1 -o gpurun_out/ncu_$name -f \
    python tools/group_mode_bench.py $key $mlog $cfg --once > gpurun_out/ncu_$name.log 2>&1
ncu -i gpurun_out/ncu_$name.ncu-rep --page raw --csv > gpurun_out/ncu_${name}_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/ncu_${name}_source.csv.gz
rm -f gpurun_out/ncu_$name.ncu-rep
python tools/ncu_summary.py gpurun_out/ncu_${name}_raw.csv gpurun_out/ncu_${name}_summary.json
