#!/bin/bash
# GPU job: test subset after the first-run fixes + ncu launch list and one full capture of the hybrid step kernel
timeout 900 python -m pytest tests/test_gpu_branches.py tests/test_gpu_config2.py tests/test_gpu_properties.py tests/test_gpu_trace.py tests/test_gpu_parity.py -m gpu -q --timeout=900 -p no:cacheprovider -x -k "not stress" 2>&1 | tail -60 > gpurun_out/t2.log
tail -5 gpurun_out/t2.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config4 --no-bullet-order --e2e-steps 2"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_h1_launches.csv $B > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:snk_hyb_step_kernel --launch-skip 4 -c 1 -f -o gpurun_out/r02_h1_full $B > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out | tail -8
