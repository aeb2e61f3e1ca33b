#!/bin/bash
# GPU job 38: stress of the split hand-out protocol; the float64 host path at the per-rank batch of the 8-GPU run with 2 / 4 / 16 host threads
timeout 900 python -m pytest tests/test_gpu_properties.py tests/test_gpu_parity.py -m gpu -q -x -k "split_hand_out or numpy_path" -p no:cacheprovider 2>&1 | tail -3
for th in 2 4 16; do SNK_HOST_THREADS=$th timeout 300 python bench.py --envs 131072 --steps 10 --warmup 3 --no-cpu-baseline --no-config4 --no-bullet-order > gpurun_out/b38_$th.log 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/b38_$th.log').read().strip().splitlines()[-1]); print('threads=$th value', round(d['value']), 'e2e', round(d['e2e']['value']), 'pinned', round(d['e2e']['value_pinned_f32']))"; done
