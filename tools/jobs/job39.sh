#!/bin/bash
# GPU job 39: evict-last stores of reward / done / ticks (DRAM traffic of the step kernel), float4 action reads of the predict kernel:
# parity tests, mapped-memory breakdown, launch list + full capture
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -q -x -k "edge_cases or env_steps_vs_oracle or numpy_path or hand_out_policy or split_hand_out_does" -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/bench_zero_copy.py 2>&1 | tail -4
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config4 --no-bullet-order --e2e-steps 2"
timeout 300 $B > gpurun_out/b39_short.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_h5_launches.csv $B > gpurun_out/ncu_l5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:snk_hyb_step_kernel --launch-skip 4 -c 1 -f -o gpurun_out/r02_h5_full $B > gpurun_out/ncu_f5.log 2>&1
tail -1 gpurun_out/b39_short.log | cut -c1-200
