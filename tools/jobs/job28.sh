#!/bin/bash
# GPU job 28: why does the manifold leg of bench.py run slower than tools/bench_manifold.py?
for i in 1 2; do timeout 300 python tools/bench_manifold.py --envs 262144 2>&1 | tail -1 | cut -c1-120; done
timeout 700 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config4 --e2e-steps 2 > gpurun_out/b28.log 2> gpurun_out/b28.err; tail -1 gpurun_out/b28.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',round(d['value'])); print('manifold', d['manifold']['ms_per_step'], 'bullet', d['bullet_order']['ms_per_step'])"
timeout 300 python tools/bench_manifold.py --envs 262144 --steps 6 2>&1 | tail -1 | cut -c1-120
