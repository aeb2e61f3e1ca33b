#!/bin/bash
# GPU job 5: PPO caller eager vs CUDA graph with breakdown; bench after the hand-out rule; alignment test
timeout 300 python tools/bench_callers.py ppo > gpurun_out/ppo_eager.log 2>&1; tail -1 gpurun_out/ppo_eager.log
timeout 300 python tools/bench_callers.py ppo --graph > gpurun_out/ppo_graph.log 2>&1; tail -2 gpurun_out/ppo_graph.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k edge_cases -p no:cacheprovider 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-bullet-order --no-config4 --e2e-steps 10 > gpurun_out/b5.log 2> gpurun_out/b5.err
python -c "import json;d=json.loads(open('gpurun_out/b5.log').read().strip().splitlines()[-1]);print('bench',round(d['value']),d['ms_per_step'],d['e2e']['value'],d['e2e']['value_pinned_f32'])"
