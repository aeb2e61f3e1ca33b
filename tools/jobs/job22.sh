#!/bin/bash
# GPU job 22: new manifold tests (host path, config-2 free running), PPO caller with TF32 policy
timeout 1200 python -m pytest tests/test_gpu_manifold.py tests/test_gpu_config2.py -m gpu -q --timeout=1000 -p no:cacheprovider -s -k "manifold" 2>&1 | tail -25 > gpurun_out/t22.log
tail -22 gpurun_out/t22.log
for o in "--graph --no-validate" "--graph --no-validate --tf32"; do timeout 300 python tools/bench_callers.py ppo $o > gpurun_out/ppo22.log 2>&1; tail -1 gpurun_out/ppo22.log | cut -c1-700; done
