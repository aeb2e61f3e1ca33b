#!/bin/bash
# GPU job 12: block Gauss-Seidel in the Bullet-order kernel: GPU suite + timing
timeout 200 python tools/bench_bullet_order.py --envs 32768 --steps 2 > gpurun_out/bo12.log 2>&1; tail -1 gpurun_out/bo12.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=1200 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/t12.log
tail -25 gpurun_out/t12.log
timeout 300 python tests/config1_gait.py > gpurun_out/gait12.log 2>&1; tail -5 gpurun_out/gait12.log
