#!/bin/bash
# GPU job 27: default bench line after the side-leg refactoring
timeout 700 python bench.py > gpurun_out/b27.log 2> gpurun_out/b27.err; tail -1 gpurun_out/b27.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'])
for k in ('hbm_bound_kernels','bullet_order','manifold'): print(k, json.dumps(d.get(k))[:400])
print('cpu',d.get('cpu_baseline',{}).get('value'))"; tail -2 gpurun_out/b27.err
