#!/bin/bash
# GPU job 16: manifold kernel with prefetching solver loops: tests + timing
timeout 900 python -m pytest tests/test_gpu_manifold.py -m gpu -q --timeout=800 -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/bench_manifold.py > gpurun_out/man16.log 2>&1
tail -3 gpurun_out/man16.log
