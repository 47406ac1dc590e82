python -m pytest tests -q -m gpu -x -k "mle or covariance or band or gram or fullsize" 2>&1 | tail -6 > gpurun_out/r2b_gram_tests.log
python tools/bench_aux.py 24 > gpurun_out/r2b_aux_dmma.json 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_mma_kernel -c 1 -o gpurun_out/r2b_gram python tools/bench_aux.py 22 > gpurun_out/r2b_gram_ncu.log 2>&1
ncu -i gpurun_out/r2b_gram.ncu-rep --page raw --csv > gpurun_out/r2b_gram_raw.csv 2>/dev/null
ncu -i gpurun_out/r2b_gram.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/r2b_gram_source.csv.gz
rm -f gpurun_out/r2b_gram.ncu-rep
