#!/bin/bash
# GPU job 30: memcheck of the two new kernels at small sizes (if the tool is allowed on this pool)
which compute-sanitizer || ls /usr/local/cuda/bin | grep -i sanit
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/bench_manifold.py --envs 2048 --steps 1 > gpurun_out/san_man.log 2>&1; echo "manifold rc=$?"; tail -4 gpurun_out/san_man.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/bench_bullet_order.py --envs 512 --steps 1 > gpurun_out/san_bo.log 2>&1; echo "bullet-order rc=$?"; tail -4 gpurun_out/san_bo.log
