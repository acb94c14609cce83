# source-level ncu capture of chol_block_kernel (the serial chain of the device fit), finest sampling interval; CSV only comes back
name=chol_block
ncu --set full --sampling-interval 0 --clock-control none --import-source on -k regex:chol_block_kernel -s 8 -c 1 -o gpurun_out/ncu_$name -f python tools/profile_aux.py > gpurun_out/ncu_$name.log 2>&1
ncu -i gpurun_out/ncu_$name.ncu-rep --page raw --csv > gpurun_out/ncu_${name}_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/ncu_${name}_source.csv.gz
rm -f gpurun_out/ncu_$name.ncu-rep
