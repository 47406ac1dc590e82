#!/bin/bash
# 1 -> N GPU strong-scaling run of bench.py on one box (what the driver does at round end).
set -u
N=${1:-8}
mkdir -p gpurun_out
for g in 1 2 4 8; do
  [ $g -gt $N ] && break
  if [ $g -eq 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/scale_${g}.json 2> gpurun_out/scale_${g}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29600 + g)) \
      bench.py --gpus $g --steps 20 --warmup 3 > gpurun_out/scale_${g}.json 2> gpurun_out/scale_${g}.err
  fi
  tail -n 1 gpurun_out/scale_${g}.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'GPUs', round(d['value'],1), 'it/s', round(d['ms_per_step'],3), 'ms/step', 'e2e', round(d['e2e']['value'],1), d['kernel_ms'])"
done
