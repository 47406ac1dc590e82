nvidia-smi -L | wc -l > gpurun_out/r2d_multi_ngpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "2-3-softplus or 8-1-softplus or 2-1-softplus-none" 2>&1 | tail -8 > gpurun_out/r2d_pytest_multi8.log
for g in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29600 + g)) bench.py --gpus $g --steps 20 --warmup 3 > gpurun_out/r2d_scale_${g}.json 2> gpurun_out/r2d_scale_${g}.err
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2d_scale_1.json 2> gpurun_out/r2d_scale_1.err
