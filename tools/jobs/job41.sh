#!/bin/bash
# GPU job 41: the default bench line with the config-3 leg
S=$(date +%s); timeout 900 python bench.py > gpurun_out/b41.log 2> gpurun_out/b41.err; echo "rc=$? wall=$(( $(date +%s) - S )) s"
python - <<'P'
import json
d=json.loads(open('gpurun_out/b41.log').read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'])
for k in ('config3_ppo_rollout','config4_ars_sweep','bullet_order','manifold','hbm_bound_kernels'):
    print(k, json.dumps(d.get(k))[:400])
print('roofline', d['roofline']['traffic'], d['roofline']['traffic_over_algorithmic'], d['roofline']['issue']['frac'])
print(d['cpu_baseline'])
P
tail -3 gpurun_out/b41.err
