#!/bin/bash
# 8-GPU job: bench at N=8 (strong scaling of 2^20 environments, collective block, config 4: 262 144 environments + all-gather)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/b_n8.log 2> gpurun_out/b_n8.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/b_n8.log').read().strip().splitlines()[-1])
print('N=8', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), round(d['e2e']['value_pinned_f32']))
print(d.get('collective')); print(d.get('config4_ars_sweep'))
P
tail -3 gpurun_out/b_n8.err
