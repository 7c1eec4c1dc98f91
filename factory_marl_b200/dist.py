"""Multi-GPU plumbing: env shards are independent per rank (the reference runs every env in its own process,
learning.py:98-100); the only collective of the path is the PPO gradient all-reduce.  torch.distributed only."""
from __future__ import annotations

import torch
import torch.distributed as dist


def env_shard(total_envs: int, rank: int, world: int):
    """Contiguous env range [lo, hi) owned by `rank` (BASELINE.json configs[4]: 65536 envs over 1/2/4/8 GPUs)."""
    base, rem = divmod(total_envs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_seeds(base_seed: int, total_envs: int, rank: int, world: int):
    lo, hi = env_shard(total_envs, rank, world)
    return list(range(base_seed + lo, base_seed + hi))


def allreduce_sum_(flat: torch.Tensor) -> torch.Tensor:
    """THE collective of the path: one sum all-reduce of the flat fp32 gradient per optimiser step (NCCL over NVLink on the GPU
    box, gloo in the CPU tests).  The division by the world size happens inside the clip + Adam kernel (`grad_scale`), because the
    global-norm clip must see the averaged gradient (ppo.PPO.train)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def allreduce_mean_(flat: torch.Tensor) -> torch.Tensor:
    allreduce_sum_(flat)
    if dist.is_initialized() and dist.get_world_size() > 1:
        flat.div_(dist.get_world_size())
    return flat


def max_over_ranks(x: float, device=None) -> float:
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
