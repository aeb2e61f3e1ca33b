#!/bin/bash
# GPU job 23: the HBM-bound helper kernels against the measured copy bandwidth
timeout 300 python tools/bench_hbm_kernels.py > gpurun_out/hbm23.log 2>&1; cat gpurun_out/hbm23.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
for k in d['kernels']: print('%-40s %8.3f ms %8.1f GB/s  %.3f of peak'%(k['kernel'],k['ms'],k['GB/s'],k['frac_of_measured_hbm_peak']))" || tail -5 gpurun_out/hbm23.log
