"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: env sharding and the gradient averaging contract."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from factory_marl_b200.dist import allreduce_mean_, env_shard, max_over_ranks, shard_seeds


def test_env_shard_partition():
    for total, world in ((65536, 8), (4096, 2), (10, 4), (7, 8)):
        spans = [env_shard(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert shard_seeds(42, 8, 1, 2) == [46, 47, 48, 49]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.full((1000,), float(rank + 1))
    allreduce_mean_(g)
    m = max_over_ranks(10.0 + rank)
    lo, hi = env_shard(4096, rank, world)
    q.put((rank, float(g[0]), m, lo, hi))
    dist.destroy_process_group()


def test_gradient_mean_and_max_timing_over_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    assert out[0][1] == out[1][1] == 1.5            # (1 + 2) / 2 on both ranks
    assert out[0][2] == out[1][2] == 11.0           # max over ranks
    assert (out[0][3], out[0][4], out[1][3], out[1][4]) == (0, 2048, 2048, 4096)
