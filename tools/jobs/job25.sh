#!/bin/bash
# GPU job 25: reset kernel with aligned write-back: HBM figures + whole GPU suite
timeout 300 python tools/bench_hbm_kernels.py > gpurun_out/hbm25.log 2>&1; python -c "
import sys,json
d=json.loads(open('gpurun_out/hbm25.log').read())
for k in d['kernels']: print('%-40s %8.3f ms %8.1f GB/s  %.3f of peak'%(k['kernel'],k['ms'],k['GB/s'],k['frac_of_measured_hbm_peak']))" || tail -5 gpurun_out/hbm25.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=1200 -p no:cacheprovider 2>&1 | tail -4
