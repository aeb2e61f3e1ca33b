"""Smallest case that exercises every kernel (for compute-sanitizer): both solvers, reset/observe/tick/step, host path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bullet_envs_b200 import SnakeVecEnv, default_params, gait_params
rng = np.random.default_rng(0)
for params in (default_params(), default_params(motor_solver=0), gait_params()):
    n = 300 if params.motor_solver == 2 and np.isinf(params.motor_max_force) else 12
    env = SnakeVecEnv(num_envs=n, device=0, params=params)
    env.reset()
    for t in range(2):
        env.step(rng.uniform(-1, 1, (n, 8)).astype(np.float32))
        env.step(torch.rand((n, 8), device="cuda") * 2 - 1)
    env.tick(np.zeros((n, 16), np.float32), 2)
    env.observe(); s = env.get_state(); env.set_state(s); env.reset(mask=np.arange(n) % 2)
    print("ok", env.counters()); env.close()
