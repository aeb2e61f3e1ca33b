#!/bin/bash
# GPU job 40: PPO rollout collection (config 3, 65 536 envs = 1.73 waves) with the split hand-out, against SNK_EXACT_SPLIT=0
for sp in 1 0; do
  SNK_EXACT_SPLIT=$sp timeout 200 python tools/bench_sizes.py 65536 2>&1 | tail -1
  for o in "--graph --no-validate" "--graph --no-validate --tf32"; do SNK_EXACT_SPLIT=$sp timeout 300 python tools/bench_callers.py ppo $o > gpurun_out/ppo40_$sp.log 2>&1; tail -1 gpurun_out/ppo40_$sp.log | cut -c1-600; done
done
