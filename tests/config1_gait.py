#!/usr/bin/env python3
"""(Test-side script: it uses the CPU oracle, so it lives under tests/.)  BASELINE.json configs[0]: the open-loop sinusoidal gait of snake_gait_test.py for 1000 ticks on a single
snake -- dt = 0.01, g = -9.81, 4 N.m motors, targets theta_n(t) = -(pi/6) sin(4 n + 2 t) on the odd joints with
t = tick * 0.01 (the script's wall clock made deterministic), obstacle block omitted (snake_gait_test.py:50-53,
64-104).  Runs the CPU oracle (fp64) and the CUDA path (N = 1 and N = 4096 replicas) and prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from bullet_envs_b200 import SnakeVecEnv, gait_params
from oracle.oracle_py import Oracle

TICKS = 1000
nn = np.arange(16)
targets = np.stack([np.where(nn % 2 == 1, -(np.pi / 6) * np.sin(4 * nn + 2 * k * 0.01), 0.0) for k in range(TICKS)]).astype(np.float32)

p = gait_params()
o = Oracle(1, p); o.reset()
t0 = time.perf_counter()
qo = []
for k in range(TICKS):
    o.tick(targets[k][None].astype(np.float64), 1)
    qo.append(o.get_state()[0].copy())
cpu_s = time.perf_counter() - t0
qo = np.array(qo)

out = {"config": "single snake, sinusoidal gait, 1000 ticks, dt 0.01, 4 N.m", "cpu_oracle_ticks_per_s": TICKS / cpu_s,
       "oracle_final_base_xy": qo[-1, 0:2].tolist()}
for n in (1, 4096):
    env = SnakeVecEnv(num_envs=n, device=0, params=p)
    env.reset(as_torch=True)
    tg = torch.from_numpy(targets).cuda()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    traj = []
    for k in range(TICKS):
        env.tick(tg[k][None].expand(n, 16).contiguous(), 1)
        if n == 1 or k == TICKS - 1:
            traj.append(env.get_state()[0].cpu().numpy())
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if n == 1:
        tr = np.array(traj)
        err = np.abs(tr[:, 13:29] - qo[:, 13:29]).max(1)
        out["gpu_n1"] = {"ticks_per_s": TICKS / dt, "q_err_max_first_100_ticks": float(err[:100].max()), "q_err_max_1000_ticks": float(err.max()),
                         "final_base_xy": tr[-1, 0:2].tolist(), "base_xy_err_final": float(np.abs(tr[-1, 0:2] - qo[-1, 0:2]).max())}
    else:
        s = env.get_state().cpu().numpy()
        out["gpu_n4096"] = {"env_ticks_per_s": n * TICKS / dt, "replicas_bit_identical": bool((s == s[0]).all())}
    env.close()
print(json.dumps(out))
