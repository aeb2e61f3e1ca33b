#!/bin/bash
# GPU job 42: streaming stores in the widening of snk_step_host_f64 (SNK_HOST_NT=1, default) against plain stores, with two host threads at
# the per-rank batch of the 8-GPU run and with the default pool at 2^20
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "numpy_path" -p no:cacheprovider 2>&1 | tail -2
for nt in 1 0 1 0; do SNK_HOST_NT=$nt SNK_HOST_THREADS=2 timeout 300 python bench.py --envs 131072 --steps 5 --warmup 3 --e2e-steps 20 --no-cpu-baseline --no-config4 --no-bullet-order > gpurun_out/b42_$nt.log 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/b42_$nt.log').read().strip().splitlines()[-1]); print('131072 envs, 2 threads, nt=$nt value', round(d['value']), 'e2e', round(d['e2e']['value']), 'pinned', round(d['e2e']['value_pinned_f32']))"; done
for nt in 1 0; do SNK_HOST_NT=$nt timeout 300 python bench.py --steps 5 --warmup 3 --e2e-steps 10 --no-cpu-baseline --no-config4 --no-bullet-order > gpurun_out/b42f_$nt.log 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/b42f_$nt.log').read().strip().splitlines()[-1]); print('2^20 envs, nt=$nt value', round(d['value']), 'e2e', round(d['e2e']['value']), 'pinned', round(d['e2e']['value_pinned_f32']))"; done
