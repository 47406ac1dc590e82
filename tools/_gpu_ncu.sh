set -u
exp() {  # keep the CSV pages, drop the (large) report
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/$1_source.csv.gz
  rm -f gpurun_out/$1.ncu-rep
}
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sweep|reduce|update|fit_small|project" -c 120 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_staged_kernel -s 4 -c 1 -o gpurun_out/r2_prof_sweep python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_ncu_f.log 2>&1
exp r2_prof_sweep
python tools/bench_configs.py 4 > gpurun_out/r2_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sweep_subunit_kernel -s 6 -c 1 -o gpurun_out/r2_prof_sub python tools/bench_configs.py 4 > gpurun_out/r2_ncu_s.log 2>&1
exp r2_prof_sub
python tools/bench_configs.py 1 > gpurun_out/r2_plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fit_small_kernel -s 1 -c 1 -o gpurun_out/r2_prof_small python tools/bench_configs.py 1 > gpurun_out/r2_ncu_p.log 2>&1
exp r2_prof_small
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sweep|reduce|update" -s 20 -c 24 --csv --log-file gpurun_out/r2_launches_config4.csv python tools/bench_configs.py 4 > gpurun_out/r2_ncu_l4.log 2>&1
python tools/time_fit_phases.py 25 20 > gpurun_out/r2_fit_phases.log 2>&1
du -sh gpurun_out > gpurun_out/r2_du.txt
