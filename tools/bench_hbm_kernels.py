"""The HBM-bound kernels of the path against the measured copy bandwidth (MEASURED_PEAKS.json): the fused env-step keeps an
environment on chip for ~30 ticks, so HBM is the right roofline only for the kernels that touch every byte once:

  snk_observe      Snake.getObservation for the whole batch: 256 B state record in, 224 B observation row out
  snk_reset        masked soft reset + observation: the same traffic plus the record written back
  snk_gae          compute_gae over a [T, N] rollout: 9 B in, 4 B (+4 with advantages) out per (t, env)
  snk_get_state    device-to-device copy of the state array (the copy-bandwidth yardstick itself: 256 B in, 256 B out)

Each is timed with CUDA events over `--iters` back-to-back launches on buffers larger than L2 (2^20 environments).

    python tools/bench_hbm_kernels.py [--envs 1048576] [--iters 20]
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bullet_envs_b200 import SnakeVecEnv, _abi  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def measure(env, iters=20, with_reset=True, peak=None):
    """GB/s of the HBM-bound kernels on `env`'s batch (a SnakeVecEnv on the current device).  with_reset=False leaves the state untouched."""
    if peak is None:
        peak = 6549.8
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
    n = env.num_envs
    lib, h = env._lib, env._h
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    dev = torch.device("cuda", torch.cuda.current_device())
    obs = torch.empty((n, 56), device=dev)
    state = torch.empty((n, 64), device=dev)
    out = []

    def row(name, ms, nbytes, what):
        gbs = nbytes / (ms * 1e-3) / 1e9
        out.append({"kernel": name, "ms": ms, "bytes": nbytes, "GB/s": gbs, "frac_of_measured_hbm_peak": gbs / peak, "traffic": what})

    row("snk_observe", timed(lambda: _abi.check(lib.snk_observe(h, p(obs), st()), lib), iters), n * (256 + 224),
        "256 B record in, 224 B observation row out per environment")
    if with_reset:
        g = torch.Generator(device=dev).manual_seed(0)
        mask = (torch.rand(n, device=dev, generator=g) < 0.5).to(torch.uint8)
        row("snk_reset (half of the batch masked)", timed(lambda: _abi.check(lib.snk_reset(h, p(mask), p(obs), st()), lib), iters), n * (256 + 224 + 128 + 1),
            "256 B in, 224 B out, the record of every masked environment (half of them) written back, 1 B mask")
    row("snk_get_state", timed(lambda: _abi.check(lib.snk_get_state(h, p(state), st()), lib), iters), n * 512, "256 B in, 256 B out")
    T = 20
    rewards = torch.rand((T, n), device=dev) * 2 - 1
    values = torch.rand((T, n), device=dev) * 2 - 1
    dones = (torch.rand((T, n), device=dev) < 0.01).to(torch.uint8)
    nv = torch.zeros(n, device=dev)
    returns = torch.empty_like(rewards); adv = torch.empty_like(rewards)
    row("snk_gae (T = 20)", timed(lambda: _abi.check(lib.snk_gae(dev.index, p(rewards), p(dones), p(values), p(nv), 0.99, 0.95, p(returns), p(adv), T, n, st()), lib),
                                  iters), T * n * (4 + 4 + 1 + 4 + 4) + n * 4, "r 4 + V 4 + done 1 in, return 4 + advantage 4 out per (t, env)")
    return {"envs": n, "hbm_peak_GB/s": peak, "kernels": out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    env = SnakeVecEnv(num_envs=args.envs, device=0)
    env.reset(as_torch=True)
    g = torch.Generator(device="cuda").manual_seed(0)
    env.step(torch.rand((args.envs, 8), device="cuda", generator=g) * 2 - 1)  # a non-trivial state
    res = measure(env, args.iters)
    env.close()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
