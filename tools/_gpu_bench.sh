( time python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err ) 2> gpurun_out/r2b_bench_time.txt
( time python bench.py --impl reference > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err ) 2> gpurun_out/r2b_bench_ref_time.txt
( time python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1 ) 2> gpurun_out/r2b_smoke_time.txt
