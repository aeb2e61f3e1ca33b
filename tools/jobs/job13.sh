#!/bin/bash
# GPU job 13: ncu full capture of the block-GS Bullet-order kernel
timeout 400 ncu --set full --clock-control none --import-source on -k regex:snk_env_kernel --launch-skip 1 -c 1 -f -o gpurun_out/r02_bo2_full python tools/bench_bullet_order.py --envs 32768 --steps 1 > gpurun_out/ncu_bo2.log 2>&1
ls -la gpurun_out | tail -3
