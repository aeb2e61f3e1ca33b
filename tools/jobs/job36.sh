#!/bin/bash
# GPU job 36: split hand-out as a separate instantiation: bit-identity, the hand-out / host-path tests, sizes
timeout 900 python -m pytest tests/test_gpu_properties.py tests/test_gpu_parity.py -m gpu -q -x -k "split_hand_out or hand_out_policy or numpy_path or host_path" -p no:cacheprovider 2>&1 | tail -3
timeout 600 python tools/bench_sizes.py 131072 1048576 1048576 2>&1 | tail -3
