timeout 900 python -m pytest tests/test_gpu_stress.py -q -m gpu -x 2>&1 | tail -30 > gpurun_out/r2_stress.log
timeout 300 python -m pytest tests/test_gpu_fit.py -q -m gpu -k "persistent or readme" 2>&1 | tail -5 > gpurun_out/r2_small_tests.log
for g in 0 64; do RFB_FIT_SMALL_CTAS=$g timeout 120 python tools/bench_configs.py 1 2>> gpurun_out/r2_config1.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$g', d['iterations_per_s'], d['ms_per_iter']*1e3,'us')" >> gpurun_out/r2_config1_ctas4.txt; done
