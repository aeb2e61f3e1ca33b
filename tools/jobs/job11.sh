#!/bin/bash
# GPU job 11: Bullet-order (warp-per-env) kernel alone: timing + one ncu full capture
timeout 200 python tools/bench_bullet_order.py --envs 32768 --steps 2 > gpurun_out/bo11.log 2>&1; tail -1 gpurun_out/bo11.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:snk_env_kernel --launch-skip 1 -c 1 -f -o gpurun_out/r02_bo_full python tools/bench_bullet_order.py --envs 32768 --steps 1 > gpurun_out/ncu_bo.log 2>&1
ls -la gpurun_out | tail -4
