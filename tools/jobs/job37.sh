#!/bin/bash
# GPU job 37: ncu launch list + full capture of the step kernel of the final build (split hand-out in its own instantiation)
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config4 --no-bullet-order --e2e-steps 2"
timeout 300 $B > gpurun_out/b37_short.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_h4_launches.csv $B > gpurun_out/ncu_l4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:snk_hyb_step_kernel --launch-skip 4 -c 1 -f -o gpurun_out/r02_h4_full $B > gpurun_out/ncu_f4.log 2>&1
tail -1 gpurun_out/b37_short.log | cut -c1-300
