python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r2f_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sweep|reduce|update|fit_small|project|gram" -c 200 --csv --log-file gpurun_out/r2f_launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_ncu_smoke.log 2>&1
