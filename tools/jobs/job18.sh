#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_manifold.py -m gpu -q --timeout=800 -p no:cacheprovider 2>&1 | tail -2
timeout 300 python tools/bench_manifold.py --envs 65536,262144 > gpurun_out/man18.log 2>&1
tail -2 gpurun_out/man18.log
