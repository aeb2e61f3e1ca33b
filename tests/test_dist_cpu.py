"""world_size-2 gloo tests of the multi-rank host logic (sharding, return gather, Welford merge)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bullet_envs_b200.dist import shard_range


def test_shard_range_partitions():
    for total in (0, 1, 7, 4096, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            parts = [shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bullet_envs_b200 import dist as sd
    lo, hi = sd.shard_range(total, rank, world)
    # returns keyed by GLOBAL env id so the gathered vector is independent of the world size
    glob = torch.arange(total, dtype=torch.float32) * 0.5 - 3.0
    full = sd.gather_returns(glob[lo:hi].clone(), total)
    ok_gather = torch.equal(full, glob)
    # Welford merge of per-shard statistics equals the statistics of the whole batch
    g = torch.Generator().manual_seed(0)
    data = torch.randn((total, 5), generator=g, dtype=torch.float64)
    part = data[lo:hi]
    cnt = torch.tensor(float(part.shape[0]), dtype=torch.float64)
    mean = part.mean(0) if part.shape[0] else torch.zeros(5, dtype=torch.float64)
    m2 = ((part - mean) ** 2).sum(0)
    n, gm, gm2 = sd.merge_welford(cnt, mean, m2)
    ok_w = bool(abs(float(n) - total) < 1e-9 and torch.allclose(gm, data.mean(0), atol=1e-12)
                and torch.allclose(gm2, ((data - data.mean(0)) ** 2).sum(0), atol=1e-9))
    ms = sd.mean_scalar(torch.tensor(float(rank)))
    q.put((rank, ok_gather, ok_w, float(ms)))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [64, 37])
def test_gather_and_welford_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, okg, okw, ms in res:
        assert okg and okw and ms == pytest.approx(0.5)


def test_global_streams_do_not_depend_on_the_sharding():
    """actions / directions keyed by the global environment index: a rank's slice is the slice of the full stream."""
    import torch
    from bullet_envs_b200.dist import global_normal, global_uniform, shard_range
    total = 1003
    full_u, full_n = global_uniform(3, 7, 0, total, 8), global_normal(3, 7, 0, total, 5)
    for world in (2, 8):
        parts_u = [global_uniform(3, 7, *shard_range(total, r, world), 8) for r in range(world)]
        parts_n = [global_normal(3, 7, *shard_range(total, r, world), 5) for r in range(world)]
        assert torch.equal(torch.cat(parts_u), full_u) and torch.equal(torch.cat(parts_n), full_n)
    assert full_u.min() >= -1 and full_u.max() < 1 and abs(float(full_u.mean())) < 0.05 and abs(float(full_u.std()) - 3 ** -0.5) < 0.02
    assert not torch.equal(global_uniform(3, 8, 0, total, 8), full_u) and not torch.equal(global_uniform(4, 7, 0, total, 8), full_u)
    assert abs(float(full_n.mean())) < 0.1 and abs(float(full_n.std()) - 1) < 0.1
