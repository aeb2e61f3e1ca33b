#!/bin/bash
# GPU job 24: reset/observe kernel with 16 threads per environment; GAE with 10 steps in flight; tests that use them
for u in 4 10; do SNK_GAE_UN=$u timeout 300 python tools/bench_hbm_kernels.py > gpurun_out/hbm24_$u.log 2>&1; python -c "
import sys,json
d=json.loads(open('gpurun_out/hbm24_$u.log').read())
for k in d['kernels']: print('UN=$u %-40s %8.3f ms %8.1f GB/s  %.3f of peak'%(k['kernel'],k['ms'],k['GB/s'],k['frac_of_measured_hbm_peak']))" || tail -5 gpurun_out/hbm24_$u.log; done
timeout 1500 python -m pytest tests -m gpu -q --timeout=1200 -p no:cacheprovider 2>&1 | tail -5
