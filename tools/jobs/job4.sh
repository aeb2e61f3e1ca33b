#!/bin/bash
# GPU job 4: balanced two-pool hand-out: invariance tests, batch-size sweep with and without it, bench
timeout 600 python -m pytest tests/test_gpu_properties.py tests/test_gpu_parity.py -m gpu -q --timeout=600 -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/t4.log
tail -4 gpurun_out/t4.log
timeout 300 python tools/bench_sizes.py 37888 50000 65536 100000 131072 200000 262144 524288 > gpurun_out/sizes_bal1.log 2>&1; cat gpurun_out/sizes_bal1.log
SNK_EXACT_BALANCE=0 timeout 300 python tools/bench_sizes.py 37888 50000 65536 100000 131072 200000 262144 524288 > gpurun_out/sizes_bal0.log 2>&1; cat gpurun_out/sizes_bal0.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-bullet-order --e2e-steps 5 > gpurun_out/b4.log 2> gpurun_out/b4.err
python -c "import json;d=json.loads(open('gpurun_out/b4.log').read().strip().splitlines()[-1]);print('bench',round(d['value']),d['ms_per_step'],d['e2e']['value'],d['e2e']['value_pinned_f32'],d['config4_ars_sweep']['env_steps_per_s'])"
