#!/bin/bash
# GPU job 29: manifold kernel: ring depth / CTAs per SM / carveout
V=$PWD/bullet_envs_b200/csrc/variants
run() { timeout 300 python tools/bench_manifold.py --envs 262144 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],1), 'ms', round(d['env_steps_per_s']))"; }
run "default(2x6)"
SNK_MAN_CARVEOUT=100 run "2x6 carveout100"
SNK_MAN_CARVEOUT=40 run "2x6 carveout40"
SNK_B200_LIB=$V/libsnake_b200_r4.so run "2x4"
SNK_B200_LIB=$V/libsnake_b200_r8.so run "2x8"
SNK_B200_LIB=$V/libsnake_b200_m3r4.so run "3x4"
SNK_B200_LIB=$V/libsnake_b200_m3r4.so SNK_MAN_CARVEOUT=100 run "3x4 carveout100"
