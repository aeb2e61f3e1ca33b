#!/bin/bash
# GPU job 34: split hand-out generalised to r > L / 2 (several second parts per lane): bit-identity, then a batch-size sweep on / off / without the 3 % gate
timeout 900 python -m pytest tests/test_gpu_properties.py -m gpu -q -x -k "split_hand_out" -p no:cacheprovider 2>&1 | tail -3
SZ="50000 62000 100000 131072 162918 200000 262144 524288 1048576"
for cfg in "SNK_EXACT_SPLIT=0" "SNK_EXACT_SPLIT=1" "SNK_EXACT_SPLIT_MINPCT=0"; do
  echo "== $cfg"; env $cfg timeout 600 python tools/bench_sizes.py $SZ 2>&1 | tail -9 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['envs'], d['ms_per_step'], round(d['env_steps_per_s']))"
done
