#!/bin/bash
# GPU job 14: manifold kernel tests + timing
timeout 900 python -m pytest tests/test_gpu_manifold.py -m gpu -q --timeout=800 -p no:cacheprovider -x 2>&1 | tail -30 > gpurun_out/t14.log
tail -30 gpurun_out/t14.log
timeout 300 python - > gpurun_out/man14.log 2>&1 <<'P'
import torch, json, sys
sys.path.insert(0, '.')
from bullet_envs_b200 import SnakeVecEnv
for n in (4096, 65536):
    env = SnakeVecEnv(num_envs=n, device=0)
    env.set_manifold(True, 0.1)
    env.reset(as_torch=True)
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = torch.rand((5, n, 8), device="cuda", generator=g) * 2 - 1
    env.step(acts[0]); env.step(acts[1]); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    tk = 0
    for t in range(2, 5):
        env.step(acts[t]); tk += env.counters()["ticks"]
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    p, k = env.manifold_stats()
    print(json.dumps({"envs": n, "ms_per_step": ms, "env_steps_per_s": n / ms * 1e3, "ticks_per_s": tk / (3 * ms) * 1e3, "points_per_tick": p / k,
                      "sweeps_per_tick": env.counters()["pgs_iterations"] / env.counters()["ticks"]}))
    env.close()
P
tail -3 gpurun_out/man14.log
