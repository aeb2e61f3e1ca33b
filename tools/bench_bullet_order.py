"""Bullet-order step kernel alone (motor_solver = 0: motor rows relaxed inside the PGS, warp per environment): env-steps/s and ticks/s.

    python tools/bench_bullet_order.py [--envs 65536] [--steps 3] [--gait]   # --gait: config 1's finite-force raw ticks instead of env-steps
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bullet_envs_b200 import SnakeVecEnv  # noqa: E402
from bullet_envs_b200._abi import default_params  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=65536)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
n = args.envs
env = SnakeVecEnv(num_envs=n, device=0, params=default_params(motor_solver=0))
env.reset(as_torch=True)
g = torch.Generator(device="cuda").manual_seed(0)
acts = torch.rand((args.steps + 1, n, 8), device="cuda", generator=g) * 2 - 1
env.step(acts[0])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tk = 0
a.record()
for t in range(args.steps):
    env.step(acts[t + 1])
    tk += int(env.counters()["ticks"])
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / args.steps
print(json.dumps({"envs": n, "ms_per_step": ms, "env_steps_per_s": n / ms * 1e3, "ticks_per_s": tk / (ms * args.steps) * 1e3,
                  "pgs_sweeps_per_tick": env.counters()["pgs_iterations"] / max(1, env.counters()["ticks"])}))
