#!/bin/bash
# GPU job 31: e2e (default float64 drop-in) with and without prefaulting the caller's output arrays
for pf in 1 0; do SNK_HOST_PREFAULT=$pf timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-config4 --no-bullet-order --e2e-steps 20 > gpurun_out/b31_$pf.log 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/b31_$pf.log').read().strip().splitlines()[-1]); print('prefault=$pf value', round(d['value']), 'e2e', round(d['e2e']['value']), 'pinned', round(d['e2e']['value_pinned_f32']))"; done
SNK_HOST_PREFAULT=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-config4 --no-bullet-order --e2e-steps 20 > gpurun_out/b31_1b.log 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/b31_1b.log').read().strip().splitlines()[-1]); print('prefault=1 again value', round(d['value']), 'e2e', round(d['e2e']['value']))"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "numpy_path or host_path" -p no:cacheprovider 2>&1 | tail -2
