"""Worker of tests/test_gpu_world_invariance.py: one rank of a W-rank run of a small sharded workload (actions, ARS directions and
state noise keyed by the GLOBAL environment index); rank 0 saves the gathered per-environment returns.  With at least W GPUs the
ranks use one GPU each and NCCL; on a one-GPU box they share cuda:0 and gather over gloo (NCCL refuses two ranks on one device)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bullet_envs_b200 import SnakeVecEnv  # noqa: E402
from bullet_envs_b200 import dist as sd  # noqa: E402


def main():
    out_path, total, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    one_gpu_each = torch.cuda.device_count() >= world
    dev_index = local if one_gpu_each else 0
    torch.cuda.set_device(dev_index)
    dev = torch.device("cuda", dev_index)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if one_gpu_each else "gloo")
    lo, hi = sd.shard_range(total, rank, world)
    n = hi - lo
    env = SnakeVecEnv(num_envs=n, device=dev_index)
    env.reset(as_torch=True)
    # (a) stepwise: 3 env-steps of global-id keyed actions
    ret_a = torch.zeros(n, device=dev)
    for t in range(3):
        obs, rew, done, _ = env.step(sd.global_uniform(5, t, lo, hi, 8, dev))
        ret_a += rew
    # (b) fused ARS rollout: directions and noise keyed by the global environment index
    env.reset(as_torch=True)
    Wenv = (0.03 * sd.global_normal(7, 0, lo, hi, 8 * 56, dev)).view(n, 8, 56).contiguous()
    noise = torch.stack([(sd.global_uniform(11, t, lo, hi, 56, dev) + 1.0) * 0.5 for t in range(T)])
    ret_b = env.rollout_linear(Wenv, T, noise=noise)
    both = torch.stack([ret_a, ret_b], 1).contiguous()       # [n, 2]
    if not one_gpu_each:
        both = both.cpu()
    full = sd.gather_returns(both.view(-1), 2 * total)        # all-gather of the returns (ars/train.py:113-116)
    if rank == 0:
        np.save(out_path, full.cpu().numpy().reshape(total, 2))
    env.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
