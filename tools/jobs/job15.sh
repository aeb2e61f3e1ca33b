#!/bin/bash
# GPU job 15: rolled block-GS + 7 warps/SM, manifold tests, whole GPU suite
timeout 200 python tools/bench_bullet_order.py --envs 32768 --steps 2 > gpurun_out/bo15.log 2>&1; tail -1 gpurun_out/bo15.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=1200 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/t15.log
tail -25 gpurun_out/t15.log
timeout 300 python tests/config1_gait.py > gpurun_out/gait15.log 2>&1; tail -2 gpurun_out/gait15.log
