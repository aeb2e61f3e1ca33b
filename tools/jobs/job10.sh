#!/bin/bash
# GPU job 10: the whole GPU suite on HEAD, the default bench line, the reference arm, smoke; then the ncu launch list and one full capture
timeout 1500 python -m pytest tests -m gpu -q --timeout=1200 -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/t10.log
tail -4 gpurun_out/t10.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/b10.log 2> gpurun_out/b10.err; tail -1 gpurun_out/b10.log | cut -c1-1500
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b10_ref.log 2> gpurun_out/b10_ref.err; tail -1 gpurun_out/b10_ref.log | cut -c1-600
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config4 --no-bullet-order --e2e-steps 2"
timeout 300 $B > gpurun_out/b10_short.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_h2_launches.csv $B > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:snk_hyb_step_kernel --launch-skip 4 -c 1 -f -o gpurun_out/r02_h2_full $B > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out | tail -8
