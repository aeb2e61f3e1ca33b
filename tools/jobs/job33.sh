#!/bin/bash
# GPU job 33: split hand-out (env-steps run in two parts): bit-identity test, then the batch sizes where it applies, split on / off and three split points
timeout 900 python -m pytest tests/test_gpu_properties.py -m gpu -q -x -k "split_hand_out" -p no:cacheprovider 2>&1 | tail -5
for cfg in "SNK_EXACT_SPLIT=0" "SNK_EXACT_SPLIT_FRAC=50" "SNK_EXACT_SPLIT_FRAC=60" "SNK_EXACT_SPLIT_FRAC=70"; do
  echo "== $cfg"; env $cfg timeout 300 python tools/bench_sizes.py 50000 90000 131072 1048576 2>&1 | tail -4
done
