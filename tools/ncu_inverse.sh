# inverse path under ncu: launch list (durations of tile_gemm_async_kernel / probe_inv_kernel at n = 2048 and n = 8192) and one
# full capture of the warm one-point probe_inv_kernel launch at n = 2048; CSV only comes back
python tools/profile_inverse.py > gpurun_out/plain_inverse.log 2>&1 || exit 1
for shape in "2048 6" "8192 20"; do
  tag=$(echo $shape | tr ' ' '_')
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_inverse_$tag.csv python tools/profile_inverse.py $shape > gpurun_out/ncu_inverse_$tag.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:probe_inv_kernel -s 4 -c 1 -o gpurun_out/ncu_probe_inv -f python tools/profile_inverse.py > gpurun_out/ncu_probe_inv.log 2>&1
ncu -i gpurun_out/ncu_probe_inv.ncu-rep --page raw --csv > gpurun_out/ncu_probe_inv_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_probe_inv.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/ncu_probe_inv_source.csv.gz
ncu --set full --clock-control none -k regex:tile_gemm_async_kernel -s 6 -c 1 -o gpurun_out/ncu_tile_gemm_async -f python tools/profile_inverse.py > gpurun_out/ncu_tile_gemm_async.log 2>&1
ncu -i gpurun_out/ncu_tile_gemm_async.ncu-rep --page raw --csv > gpurun_out/ncu_tile_gemm_async_raw.csv 2>/dev/null
rm -f gpurun_out/ncu_probe_inv.ncu-rep gpurun_out/ncu_tile_gemm_async.ncu-rep
