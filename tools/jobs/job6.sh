#!/bin/bash
# GPU job 6: PPO caller (validation off / graph), chunk-pipelined float64 host path, numpy-path tests
timeout 300 python tools/bench_callers.py ppo --no-validate > gpurun_out/ppo_noval.log 2>&1; tail -1 gpurun_out/ppo_noval.log | cut -c1-600
timeout 300 python tools/bench_callers.py ppo --graph > gpurun_out/ppo_graph.log 2>&1; tail -1 gpurun_out/ppo_graph.log | cut -c1-600
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "numpy_path or host_path or edge or golden" -p no:cacheprovider 2>&1 | tail -3
for c in 1 2 4 8; do SNK_HOST_CHUNKS=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-bullet-order --no-config4 --e2e-steps 10 > gpurun_out/b6_$c.log 2> gpurun_out/b6_$c.err
python -c "import json;d=json.loads(open('gpurun_out/b6_$c.log').read().strip().splitlines()[-1]);print('chunks $c',round(d['value']),d['e2e']['value'],d['e2e']['value_pinned_f32'])"; done
