for x in 1 2 1 2; do
  RFB_XCHG=$x timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 40 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('xchg=$x', round(d['value'],1), round(d['ms_per_step'],4), d['kernel_ms']['train_sweep'], round(d['kernel_ms']['rest_of_step']*1000,1),'us', d.get('kernel_ms_eager'))"
done > gpurun_out/r2c_xchg_cmp.txt 2>&1
