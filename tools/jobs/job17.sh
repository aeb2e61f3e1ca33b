#!/bin/bash
# GPU job 17: manifold kernel, tick-granular lock step: tests + timing + ncu
timeout 900 python -m pytest tests/test_gpu_manifold.py -m gpu -q --timeout=800 -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/bench_manifold.py --envs 65536,262144 > gpurun_out/man17.log 2>&1
tail -2 gpurun_out/man17.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:snk_man_step_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r02_man_full python tools/bench_manifold.py --envs 65536 --steps 1 > gpurun_out/ncu_man.log 2>&1
ls -la gpurun_out | tail -2
