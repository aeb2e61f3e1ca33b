#!/usr/bin/env python3
"""Where do the 2 % between the device-resident env-step and the mapped-host-memory one (snk_step_host on page-locked buffers) go?
snk_step is called with every combination of {actions, results} in device memory / in mapped page-locked host memory (the same
pointers under unified addressing), CUDA events around `steps` launches:   python tools/bench_zero_copy.py [envs]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from bullet_envs_b200 import SnakeVecEnv, _abi
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    steps, warmup = 8, 3
    g = torch.Generator(device="cuda").manual_seed(0)
    acts_d = torch.rand((steps + warmup, n, 8), generator=g, device="cuda") * 2 - 1
    acts_h = acts_d.cpu().pin_memory()

    def bufs(host):
        kw = dict(pin_memory=True) if host else dict(device="cuda")
        return (torch.empty((n, 56), dtype=torch.float32, **kw), torch.empty((n,), dtype=torch.float32, **kw),
                torch.empty((n,), dtype=torch.uint8, **kw), torch.empty((n,), dtype=torch.int32, **kw))

    for a_host in (False, True):
        for o_host in (False, True):
            env = SnakeVecEnv(num_envs=n, device=0)
            env.reset(as_torch=True)
            obs, rew, done, ticks = bufs(o_host)
            acts = acts_h if a_host else acts_d
            p = lambda t: ctypes.c_void_p(t.data_ptr())
            def step(t):
                _abi.check(env._lib.snk_step(env._h, p(acts[t]), p(obs), p(rew), p(done), p(ticks), env._stream()), env._lib)
            for t in range(warmup):
                step(t)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for t in range(warmup, warmup + steps):
                step(t)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            print(json.dumps({"envs": n, "actions": "host" if a_host else "device", "results": "host" if o_host else "device",
                              "ms_per_step": round(ms, 3), "env_steps_per_s": round(n / ms * 1e3)}), flush=True)
            env.close()


if __name__ == "__main__":
    main()
