#!/bin/bash
# GPU job 21: end-of-round validation (re-run after every later change): whole GPU suite, smoke, default bench line (with the bullet_order and manifold legs), reference arm, launch list
timeout 1500 python -m pytest tests -m gpu -q --timeout=1200 -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/t21.log
tail -3 gpurun_out/t21.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 700 python bench.py > gpurun_out/b21.log 2> gpurun_out/b21.err; tail -1 gpurun_out/b21.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'issue',d['roofline'].get('issue'))
print('bullet_order',d.get('bullet_order')); print('manifold',d.get('manifold')); print('cfg4',d.get('config4_ars_sweep',{}).get('env_steps_per_s')); print('cpu',d.get('cpu_baseline',{}).get('value'))"
timeout 300 python tools/bench_manifold.py --envs 65536,262144 > gpurun_out/man21.log 2>&1; tail -2 gpurun_out/man21.log
