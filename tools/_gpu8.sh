nvidia-smi -L | wc -l > gpurun_out/r2_multi_ngpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "4-1 or 8-1 or 2-1-softmax" 2>&1 | tail -15 > gpurun_out/r2_pytest_multi8.log
for g in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29600 + g)) bench.py --gpus $g --steps 20 --warmup 5 > gpurun_out/r2_scale_${g}.json 2> gpurun_out/r2_scale_${g}.err
done
