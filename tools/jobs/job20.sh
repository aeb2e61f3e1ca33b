#!/bin/bash
# GPU job 20: ncu full captures of the final Bullet-order (rolled block sweep) and manifold (cp.async ring) kernels, each after its plain run
timeout 200 python tools/bench_bullet_order.py --envs 32768 --steps 2 > gpurun_out/bo20.log 2>&1 && tail -1 gpurun_out/bo20.log && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:snk_env_kernel --launch-skip 1 -c 1 -f -o gpurun_out/r02_bo3_full python tools/bench_bullet_order.py --envs 32768 --steps 1 > gpurun_out/ncu_bo3.log 2>&1
timeout 200 python tools/bench_manifold.py --envs 65536 > gpurun_out/man20.log 2>&1 && tail -1 gpurun_out/man20.log && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:snk_man_step_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r02_man2_full python tools/bench_manifold.py --envs 65536 --steps 1 > gpurun_out/ncu_man2.log 2>&1
ls -la gpurun_out | tail -4
