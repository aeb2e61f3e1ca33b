#!/bin/bash
# GPU job 32: flag-driven float64 host path (one launch, rows widened while the kernel runs) against the chunk-pipelined one
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "numpy_path or host_path" -p no:cacheprovider 2>&1 | tail -3
for fl in 1 0 1; do SNK_HOST_FLAGS=$fl timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-config4 --no-bullet-order --e2e-steps 20 > gpurun_out/b32_$fl.log 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/b32_$fl.log').read().strip().splitlines()[-1]); print('flags=$fl value', round(d['value']), 'e2e', round(d['e2e']['value']), 'pinned', round(d['e2e']['value_pinned_f32']))"; done
