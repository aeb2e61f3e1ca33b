#!/usr/bin/env python3
"""bench.py -- snake env-steps/s of the batched SnakeGymEnv.step() hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

Workload (BASELINE.json configs[4], SURVEY.md 8d "config 5"): 2^20 environments in total, split
evenly over the N ranks (one process per GPU, no data-path collective -- environments are
independent), fresh U[-1,1]^8 actions every step from a per-rank Philox stream generated on the
device before the timed region, soft-reset start.  A "step" is one SubprocVecEnv.step() over the
whole batch = one launch of the fused env-step kernel.

`value`   : env-steps/s with actions and output buffers resident in HBM (CUDA events, max over ranks).
`e2e`     : the same metric through the reference-facing call with HOST buffers -- SnakeVecEnv.step(numpy)
            -> C-ABI snk_step_host: pinned H2D of the actions, the kernel, D2H of obs/reward/done/ticks.
`roofline`: the env-step kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json).  The
            kernel keeps an environment on chip for ~30 physics ticks, so its algorithmic bytes are
            tiny by construction; `issue` carries the figure that actually
            bounds it (fp32 issue slots).
`cpu_baseline`: the oracle port of the reference algorithm on this box's host cores (rank 0, N=1).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOTAL_ENVS = 1 << 20
WORKLOAD = "1M-env throughput sweep (2^20 envs total, random U[-1,1] actions, 1/240 s ticks, reference defaults)"
# algorithmic bytes per env-step of the fused kernel (SURVEY.md 8d / DESIGN.md section 6):
# read 45-float minimal state 180 + action 32; write state 180 + obs 224 + reward 4 + done 1
BYTES_PER_ENV_STEP = 621
# DRAM bytes per environment of one launch, from the ncu --set full capture of this command at 2^20 environments
# (profiles/r01_v9_exact_full_raw_1m.csv: dram__bytes_read.sum 455.35 MB + dram__bytes_write.sum 592.10 MB per launch =
# 999 B/env; in index order it was 768 B/env: with the longest-first hand-out neighbouring environments are no longer
# processed at the same time, so lines they share (32 B action rows, 4 B reward/ticks, 1 B done) move more than once.
# Either way 0.03 % of the DRAM peak.)
NCU_DRAM_BYTES_PER_ENV = (455350272 + 592098816) / float(1 << 20)
# warp-level instructions executed per environment-tick, same capture: smsp__inst_executed.sum = 1.70003e11 for
# 2^20 env-steps of 30.03 ticks each (every lane slot counts: masked / converged lanes execute too)
NCU_WARP_INST_PER_ENV_TICK = 170002985085.0 / ((1 << 20) * 30.0271)


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def cpu_leg(solver, n_envs, steps, threads):
    """Oracle port on the host cores: `steps` env-steps over n_envs environments after one warm-up step."""
    import numpy as np
    from bullet_envs_b200 import default_params
    from oracle.oracle_py import Oracle
    o = Oracle(n_envs, default_params(motor_solver=solver))
    rng = np.random.default_rng(0)
    o.reset()
    o.step(rng.uniform(-1, 1, (n_envs, 8)), threads=threads)  # leaves the rest pose
    o.counters(clear=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step(rng.uniform(-1, 1, (n_envs, 8)), threads=threads)
    dt = time.perf_counter() - t0
    c = o.counters()
    o.close()
    return n_envs * steps / dt, c["ticks"] / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np  # noqa: F401
    threads = os.cpu_count() or 1
    n = 128 * threads
    # W warm-up + K timed steps, each a bounded sample of the workload (n environments)
    from bullet_envs_b200 import default_params
    from oracle.oracle_py import Oracle
    o = Oracle(n, default_params(motor_solver=0))
    rng = np.random.default_rng(0)
    o.reset()
    for _ in range(args.warmup):
        o.step(rng.uniform(-1, 1, (n, 8)), threads=threads)
    o.counters(clear=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.step(rng.uniform(-1, 1, (n, 8)), threads=threads)
    dt = time.perf_counter() - t0
    c = o.counters()
    v = n * args.steps / dt
    sample = "%d envs x %d env-steps, fp64, Bullet-order PGS rows (motor_solver=0), %d host threads" % (n, args.steps, threads)
    line = {"impl": "reference", "metric": "snake env-steps/sec", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
            "ticks_per_s": c["ticks"] / dt,
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "PyBullet is not installable in this image; this is the C oracle port of the reference algorithm (oracle/snake_oracle.c) "
                    "without the reference's 10 ms sleep per tick (snake.py:296)"}
    print(json.dumps(line))
    return 0


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise RuntimeError("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from bullet_envs_b200 import SnakeVecEnv
    from bullet_envs_b200.dist import shard_range

    total = args.envs
    lo, hi = shard_range(total, rank, world)
    n = hi - lo
    K, W = args.steps, args.warmup
    env = SnakeVecEnv(num_envs=n, device=local, obs_dtype=np.float32, pinned_io=True)  # options of the numpy path only (e2e leg)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)  # Philox, keyed per rank
    acts = torch.rand((W + K, n, 8), device=dev, generator=gen) * 2 - 1
    obs = torch.empty((n, 56), device=dev); rew = torch.empty((n,), device=dev); done = torch.empty((n,), dtype=torch.uint8, device=dev)
    tick_sum = torch.zeros((), dtype=torch.int64, device=dev)
    env.reset(as_torch=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for t in range(W):
        env.step(acts[t], out=(obs, rew, done))
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = env.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for t in range(W, W + K):
        env.step(acts[t], out=(obs, rew, done))
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    gpu_launches = env.launch_count() - launches0
    # per-launch kernel time on the launching stream (second pass, one event pair per launch) + tick statistics
    kt = []
    for t in range(W, W + K):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.step(acts[t], out=(obs, rew, done)); b.record()
        tick_sum += env.last_ticks.sum()
        kt.append((a, b))
    torch.cuda.synchronize()
    kernel_ms = sum(a.elapsed_time(b) for a, b in kt) / len(kt)
    clocks = sampler.stop() if rank == 0 else None
    c = env.counters()
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    stats = torch.tensor([float(tick_sum.item()), float(c["pgs_iterations"]), float(c["ticks"]), float(n)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    ms = float(t_ms.item())
    value = total * K / (ms * 1e-3)
    ticks_per_step_env = stats[0].item() / (K * total)

    # ---- e2e: the reference-facing call with HOST buffers (numpy in -> numpy out through snk_step_host):
    # every step copies that step's actions H2D and obs/reward/done/ticks D2H inside the timed region
    Ke = max(1, min(K, args.e2e_steps))
    host_acts = acts[W - 1:W + Ke].cpu().numpy()  # [0] = untimed warm-up action, distinct from the first timed one

    def e2e_leg(e):
        e.step(host_acts[0])  # allocates the staging buffers outside the timed region
        barrier()
        t0 = time.perf_counter()
        for t in range(1, Ke + 1):
            o_h, r_h, d_h, _ = e.step(host_acts[t])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        assert np.isfinite(r_h).all() and o_h.shape == (n, 56)
        return total * Ke / float(t_e.item())

    e2e_value = e2e_leg(env)
    # the strict drop-in defaults (fresh pageable float64 arrays every step, as SubprocVecEnv returns them)
    strict = SnakeVecEnv(num_envs=n, device=local)
    strict.reset()
    e2e_strict = e2e_leg(strict)
    strict.close()

    if rank == 0:
        peak, how = _peaks()
        achieved = BYTES_PER_ENV_STEP * n / (kernel_ms * 1e-3) / 1e9
        # the bound that actually binds: warp-instruction issue slots (4 schedulers x SMs x SM clock)
        sm_hz = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        tick_rate_per_gpu = (total / world) * ticks_per_step_env / (kernel_ms * 1e-3)
        issue = {"bound": "issue", "achieved": NCU_WARP_INST_PER_ENV_TICK * tick_rate_per_gpu, "peak": 4.0 * sms * sm_hz, "unit": "warp-inst/s",
                 "smsp_issue_active_pct_ncu": 65.6, "fma_pipe_active_pct_ncu": 48.2, "lanes_per_instruction_ncu": 32.0,
                 "source": "instructions per env-tick from profiles/r01_v9_exact_full_raw_1m.csv (ncu, same command) x the tick rate timed here"}
        issue["frac"] = issue["achieved"] / issue["peak"]
        line = {
            "metric": "snake env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_total": total, "envs_per_gpu": n, "actions": "Philox U[-1,1], device resident, new every step",
                       "l2": "inputs larger than L2: %.0f MB of state + %.0f MB of outputs per GPU per step vs 126 MB L2" % (n * 256 / 1e6, n * 233 / 1e6),
                       "solver": "motor rows eliminated (motor_solver=auto), PGS <= 50 sweeps, residual 1e-7"},
            "ticks_per_s": value * ticks_per_step_env, "ticks_per_env_step": ticks_per_step_env,
            "pgs_sweeps_per_tick_last_step": stats[1].item() / max(1.0, stats[2].item()),
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": int(total * 8 * 4),
                    "d2h_bytes_per_step": int(total * (56 * 4 + 4 + 1 + 4)), "steps": Ke,
                    "call": "SnakeVecEnv(obs_dtype=float32, pinned_io=True).step(numpy) -> snk_step_host: the step's actions are copied into page-locked host memory, the kernel reads each environment's action row from it and posts obs/reward/done/ticks rows back to page-locked host memory over PCIe as environments finish (mapped host buffers, no staging copy), then a stream sync",
                    "value_strict_dropin": e2e_strict,
                    "call_strict_dropin": "SnakeVecEnv().step(numpy) -> snk_step_host_f64: float64 actions in, fresh pageable float64 obs/reward arrays out every step (the reference's dtypes), narrowed / widened by the library's host threads around the kernel on its mapped pinned buffers"},
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_DRAM_BYTES_PER_ENV * n, "traffic_source": "ncu capture at 2^20 envs, scaled per environment",
                         "peak_source": how, "kernel": "snk_exact_step_kernel<true>", "kernel_ms": kernel_ms,
                         "bytes_per_env_step": BYTES_PER_ENV_STEP,
                         "issue": issue,
                         "note": "not HBM bound by construction: an environment stays on chip (TMEM / shared memory) for ~30 ticks x 32 contacts x <=50 "
                                 "solver sweeps per 621 B of HBM traffic; the binding limit is fp32 issue / dependent-issue latency (profiles/)"},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            nb = 512 * threads  # ~10 s of CPU work per leg
            v0, tk0, dt0 = cpu_leg(0, nb, 2, threads)
            v2, tk2, dt2 = cpu_leg(2, 4 * nb, 2, threads)
            line["cpu_baseline"] = {"value": v0, "unit": "env-steps/s", "cores": threads, "kind": "port",
                                    "sample": "oracle fp64, Bullet-order PGS rows: %d envs x 2 env-steps (%.1f s)" % (nb, dt0),
                                    "ticks_per_s": tk0, "value_exact_solver": v2,
                                    "sample_exact_solver": "oracle fp64, motor rows eliminated (the kernel's CPU twin): %d envs x 2 env-steps (%.1f s)" % (4 * nb, dt2)}
        print(json.dumps(line))
    env.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=TOTAL_ENVS, help="total environments over all ranks")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
