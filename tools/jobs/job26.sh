#!/bin/bash
# GPU job 26: Bullet-order kernel with the normal-row Jacobians in registers, 8 (then 9) warps per SM
timeout 200 python tools/bench_bullet_order.py --envs 32768 --steps 2 > gpurun_out/bo26.log 2>&1; tail -1 gpurun_out/bo26.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trace.py tests/test_gpu_branches.py -m gpu -q --timeout=1200 -p no:cacheprovider 2>&1 | tail -4
timeout 300 python tests/config1_gait.py > gpurun_out/gait26.log 2>&1; tail -1 gpurun_out/gait26.log
