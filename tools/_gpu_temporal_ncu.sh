timeout 300 ncu --set full --clock-control none --import-source on -k regex:temporal_project_const_kernel -s 1 -c 1 -o gpurun_out/r2b_temporal python tools/bench_fused.py 22 > gpurun_out/r2b_temporal_ncu.log 2>&1
ncu -i gpurun_out/r2b_temporal.ncu-rep --page raw --csv > gpurun_out/r2b_temporal_raw.csv 2>/dev/null
ncu -i gpurun_out/r2b_temporal.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/r2b_temporal_source.csv.gz
rm -f gpurun_out/r2b_temporal.ncu-rep
