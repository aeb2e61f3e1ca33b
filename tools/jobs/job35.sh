#!/bin/bash
# GPU job 35: the whole GPU suite on HEAD (flag-driven float64 host path + split hand-out), smoke, the default bench line, the reference arm;
# then the ncu launch list of the same bench command and one full capture of the step kernel
timeout 1500 python -m pytest tests -m gpu -q --timeout=1200 -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/t35.log
tail -4 gpurun_out/t35.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/b35.log 2> gpurun_out/b35.err; tail -1 gpurun_out/b35.log | cut -c1-1800
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b35_ref.log 2> gpurun_out/b35_ref.err; tail -1 gpurun_out/b35_ref.log | cut -c1-600
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config4 --no-bullet-order --e2e-steps 2"
timeout 300 $B > gpurun_out/b35_short.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_h3_launches.csv $B > gpurun_out/ncu_l3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:snk_hyb_step_kernel --launch-skip 4 -c 1 -f -o gpurun_out/r02_h3_full $B > gpurun_out/ncu_f3.log 2>&1
ls -la gpurun_out | tail -6
