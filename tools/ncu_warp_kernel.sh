name=wk_c3
ncu --set full --clock-control none --import-source on -k regex:sweep_warp_kernel -s 3 -c 1 -o gpurun_out/ncu_$name -f python tools/small_n_bench.py > gpurun_out/ncu_$name.log 2>&1
ncu -i gpurun_out/ncu_$name.ncu-rep --page raw --csv > gpurun_out/ncu_${name}_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/ncu_${name}_source.csv.gz
rm -f gpurun_out/ncu_$name.ncu-rep
python tools/ncu_summary.py gpurun_out/ncu_${name}_raw.csv gpurun_out/ncu_${name}_summary.json
