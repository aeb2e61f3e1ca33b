#!/bin/bash
# 4-GPU job: bench at N=4 (collective block + config 4)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/b_n4.log 2> gpurun_out/b_n4.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/b_n4.log').read().strip().splitlines()[-1])
print('N=4', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), round(d['e2e']['value_pinned_f32']))
print(d.get('collective')); print(d.get('config4_ars_sweep'))
P
tail -3 gpurun_out/b_n4.err
