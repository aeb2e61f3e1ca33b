"""Step kernel with persistent contact manifolds + warm starting (snk_set_manifold) alone: env-steps/s, ticks/s, cached points per tick.

    python tools/bench_manifold.py [--envs 4096,65536,262144] [--steps 3] [--warm 0.1]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bullet_envs_b200 import SnakeVecEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", default="4096,65536,262144")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warm", type=float, default=0.1)
args = ap.parse_args()
for n in [int(x) for x in args.envs.split(",")]:
    env = SnakeVecEnv(num_envs=n, device=0)
    env.set_manifold(True, args.warm)
    env.reset(as_torch=True)
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = torch.rand((args.steps + 2, n, 8), device="cuda", generator=g) * 2 - 1
    env.step(acts[0]); env.step(acts[1])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    tk = sw = pts = 0
    for t in range(args.steps):
        env.step(acts[t + 2])
        c = env.counters(); tk += c["ticks"]; sw += c["pgs_iterations"]; pts += env.manifold_stats()[0]
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    print(json.dumps({"envs": n, "ms_per_step": ms, "env_steps_per_s": n / ms * 1e3, "ticks_per_s": tk / (ms * args.steps) * 1e3,
                      "points_per_tick": pts / tk, "sweeps_per_tick": sw / tk}))
    env.close()
