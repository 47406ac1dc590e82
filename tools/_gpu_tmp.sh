for v in 0 1; do RFB_TEMPORAL_ST_SMEM=$v timeout 120 python tools/bench_fused.py 22 2>&1 | grep "fused=0" | tail -1; done > gpurun_out/r2b_stsmem.log
RFB_TEMPORAL_ST_SMEM=1 timeout 200 python -m pytest tests/test_gpu_project.py -q -m gpu -x 2>&1 | tail -3 >> gpurun_out/r2b_stsmem.log
