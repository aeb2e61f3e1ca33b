#!/bin/bash
# GPU job 3: remaining tests (incl. the stress test once), A/B of compile-time variants, batch-size sweep, PPO caller
timeout 900 python -m pytest tests/test_gpu_config2.py tests/test_gpu_rollout.py tests/test_gpu_world_invariance.py tests/test_gpu_branches.py -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/t3.log
tail -4 gpurun_out/t3.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-config4 --no-bullet-order --e2e-steps 1"
for v in default f3 f5 f6 old; do
  if [ $v = default ]; then L=""; else L="$PWD/bullet_envs_b200/csrc/variants/libsnake_b200_$v.so"; fi
  SNK_B200_LIB=$L timeout 200 $B > gpurun_out/ab_$v.log 2> gpurun_out/ab_$v.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/ab_$v.log").read().strip().splitlines()[-1]); print("$v", round(d["value"]), d["ms_per_step"], d["roofline"]["kernel_ms"])
except Exception as e: print("$v failed", e)
P
done
for w in 7 6; do SNK_EXACT_WARPS=$w timeout 200 $B > gpurun_out/ab_w$w.log 2>&1; python -c "import json;d=json.loads(open('gpurun_out/ab_w$w.log').read().strip().splitlines()[-1]);print('warps $w',round(d['value']))"; done
timeout 300 python tools/bench_sizes.py > gpurun_out/sizes_r02.log 2>&1; tail -12 gpurun_out/sizes_r02.log
timeout 300 python tools/bench_callers.py ppo > gpurun_out/ppo_r02.log 2>&1; tail -2 gpurun_out/ppo_r02.log
