"""Multi-GPU plumbing: environments shard by contiguous index range, one process per GPU.

The physics needs no exchange (every environment is independent; in the reference each lives in its
own process, ``ppo/multiprocessing_env.py:106-111``).  The only collectives are the fan-in the
reference does over pipes:

* :func:`gather_returns`  -- all-gather of per-environment episode returns (ARS ``test_envs`` return
  vector, ``ars/train.py:113-116``; 4 B per environment, latency bound on NVLink);
* :func:`merge_welford`   -- all-reduce of the running observation statistics of the ARS normaliser
  (``ars/train.py:152-169``), merged with Chan's parallel formula;
* :func:`mean_scalar`     -- all-reduce of logged scalars (``ppo/train.py:123,183``).

Backend: NCCL on the GPUs; the same code runs on ``gloo`` in the CPU tests (tests/test_dist_cpu.py).
"""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of `total` environments owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist


def is_distributed() -> bool:
    dist = _dist()
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def gather_returns(local_returns, total: int | None = None):
    """All-gather of the per-environment returns: every rank gets the full vector ordered by global
    environment index (ragged shards allowed).  ``local_returns``: 1-D tensor on the rank's device."""
    import torch
    dist = _dist()
    if not is_distributed():
        return local_returns
    world, rank = dist.get_world_size(), dist.get_rank()
    n_local = torch.tensor([local_returns.numel()], device=local_returns.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    if all(s == m for s in sizes):
        out = torch.empty(world * m, device=local_returns.device, dtype=local_returns.dtype)
        dist.all_gather_into_tensor(out, local_returns.contiguous())
    else:
        pad = torch.zeros(m, device=local_returns.device, dtype=local_returns.dtype)
        pad[:local_returns.numel()] = local_returns
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        out = torch.cat([p[:s] for p, s in zip(parts, sizes)])
    if total is not None and out.numel() != total:
        raise RuntimeError("gathered %d returns, expected %d" % (out.numel(), total))
    return out


def merge_welford(count, mean, m2):
    """Merge per-rank (count, mean[d], M2[d]) into the global statistics (Chan et al.):
    n = sum n_r;  mean = sum n_r mean_r / n;  M2 = sum [M2_r + n_r (mean_r - mean)^2].
    All three are tensors on the rank's device; returns the merged triple (identical on every rank)."""
    import torch
    dist = _dist()
    if not is_distributed():
        return count, mean, m2
    count = count.to(mean.dtype).reshape(1)
    packed = torch.cat([count, count * mean])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    n = packed[0]
    gmean = packed[1:] / n.clamp_min(1)
    corr = m2 + count * (mean - gmean) ** 2
    dist.all_reduce(corr, op=dist.ReduceOp.SUM)
    return n, gmean, corr


def mean_scalar(x):
    """Mean over ranks of a scalar tensor (logged rewards / losses)."""
    dist = _dist()
    if not is_distributed():
        return x
    y = x.clone()
    dist.all_reduce(y, op=dist.ReduceOp.SUM)
    return y / dist.get_world_size()


def global_uniform(seed: int, step: int, lo: int, hi: int, width: int, device=None, total: int | None = None):
    """U[-1, 1) numbers for the environments [lo, hi) of a sharded batch, keyed by the GLOBAL environment index: element
    (env, k) of step `step` is a hash of (seed, step, env, k) only, so a rank's slice does not depend on how many ranks
    there are (W = 1 and W = 8 see bit-identical actions / directions).  splitmix64 finaliser on a 64-bit
    counter, evaluated with wrapping int64 tensor arithmetic on `device`; returns float32 [hi - lo, width]."""
    import torch
    env = torch.arange(lo, hi, dtype=torch.int64, device=device).unsqueeze(1)
    k = torch.arange(width, dtype=torch.int64, device=device).unsqueeze(0)
    wrap = lambda v: ((int(v) + (1 << 63)) % (1 << 64)) - (1 << 63)  # Python int -> two's-complement int64
    x = (env * width + k) + wrap((int(step) + 1) * 0x632BE59BD9B4E019 + int(seed) * 0x2545F4914F6CDD1D)
    # splitmix64 (Steele, Lea, Flood 2014); arithmetic shifts are masked to emulate logical shifts on signed int64
    x = x + (-7046029254386353131)                                   # 0x9E3779B97F4A7C15
    x = (x ^ ((x >> 30) & 0x3FFFFFFFF)) * (-4658895280553007687)     # 0xBF58476D1CE4E5B9
    x = (x ^ ((x >> 27) & 0x1FFFFFFFFF)) * (-7723592293110705685)    # 0x94D049BB133111EB
    x = x ^ ((x >> 31) & 0x1FFFFFFFF)
    u = ((x >> 40) & 0xFFFFFF).to(torch.float32) * (1.0 / 8388608.0) - 1.0   # 24 bits -> [-1, 1)
    return u


def global_normal(seed: int, step: int, lo: int, hi: int, width: int, device=None):
    """N(0, 1) numbers keyed by the global environment index (Box-Muller on two :func:`global_uniform` streams): the ARS
    directions delta_i of environment i (``ars/train.py:208-219``) do not depend on the number of ranks."""
    import math
    import torch
    u1 = 1.0 - (global_uniform(seed, 2 * step, lo, hi, width, device) + 1.0) * 0.5        # (0, 1]
    u2 = (global_uniform(seed, 2 * step + 1, lo, hi, width, device) + 1.0) * 0.5          # [0, 1)
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos((2.0 * math.pi) * u2)
