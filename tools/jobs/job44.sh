#!/bin/bash
# GPU job 44: full capture of the split instantiation at the per-GPU batch of the 8-GPU run (131 072 envs), and of the plain one with SNK_EXACT_SPLIT=0
timeout 200 python tools/bench_sizes.py 131072 > gpurun_out/s44.log 2>&1 && tail -1 gpurun_out/s44.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:snk_hyb_step_kernel --launch-skip 8 -c 1 -f -o gpurun_out/r02_h5_split_128k python tools/bench_sizes.py 131072 > gpurun_out/ncu_s44.log 2>&1
SNK_EXACT_SPLIT=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:snk_hyb_step_kernel --launch-skip 8 -c 1 -f -o gpurun_out/r02_h5_twopool_128k python tools/bench_sizes.py 131072 > gpurun_out/ncu_s44b.log 2>&1
ls -la gpurun_out/*128k*
