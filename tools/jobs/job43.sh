#!/bin/bash
# GPU job 43: final validation of HEAD: the whole GPU suite, smoke, the default bench line, the reference arm
timeout 1800 python -m pytest tests -m gpu -q --timeout=1200 -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/t43.log
tail -3 gpurun_out/t43.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/b43.log 2> gpurun_out/b43.err; tail -1 gpurun_out/b43.log | cut -c1-300
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b43_ref.log 2> gpurun_out/b43_ref.err; tail -1 gpurun_out/b43_ref.log | cut -c1-200
