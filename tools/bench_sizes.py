#!/usr/bin/env python3
"""Device-resident env-step throughput over a range of batch sizes (one GPU): fresh U[-1,1] actions every step,
CUDA events around `steps` snk_step launches after `warmup` launches.  Used for the small-batch hand-out ablation
(SNK_EXACT_SPREAD=0 restores CTA-major filling):   python tools/bench_sizes.py 16 256 4096 14208 28416 65536"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from bullet_envs_b200 import SnakeVecEnv
    sizes = [int(x) for x in sys.argv[1:]] or [16, 4096, 18944, 37888, 65536, 131072, 262144]
    steps, warmup = 20, 5
    for n in sizes:
        env = SnakeVecEnv(num_envs=n, device=0)
        env.reset(as_torch=True)
        g = torch.Generator(device="cuda").manual_seed(0)
        acts = torch.rand((steps + warmup, n, 8), generator=g, device="cuda") * 2 - 1
        obs = torch.empty((n, 56), device="cuda"); rew = torch.empty((n,), device="cuda"); done = torch.empty((n,), dtype=torch.uint8, device="cuda")
        for t in range(warmup):
            env.step(acts[t], out=(obs, rew, done))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ticks = 0
        for t in range(warmup, warmup + steps):
            env.step(acts[t], out=(obs, rew, done))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(json.dumps({"envs": n, "ms_per_step": round(ms, 4), "env_steps_per_s": round(n / ms * 1e3, 1),
                          "spread": os.environ.get("SNK_EXACT_SPREAD", "3")}), flush=True)
        env.close()


if __name__ == "__main__":
    main()
