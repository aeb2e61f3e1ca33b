#!/bin/bash
# GPU job 45: first-part share of the split hand-out at r > L / 2 (65 536 and 100 000 envs): r / L + 0, 5, 10 (default), 15 points
for fr in 73 78 83 88; do echo "65536 frac=$fr $(SNK_EXACT_SPLIT_FRAC=$fr timeout 100 python tools/bench_sizes.py 65536 2>&1 | tail -1 | cut -c1-80)"; done
for fr in 64 69 74 79; do echo "100000 frac=$fr $(SNK_EXACT_SPLIT_FRAC=$fr timeout 100 python tools/bench_sizes.py 100000 2>&1 | tail -1 | cut -c1-80)"; done
