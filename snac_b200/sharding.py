"""Env sharding over the GPUs of one box.

Envs are independent, so rank r owns one contiguous slice of the global env index space and there is no
data-path collective.  The only exchange is the episode-statistics vector
(sum of returns, sum of final IoUs, episodes, steps), summed over ranks with one all-reduce
(NCCL over NVLink on GPUs; the same code path runs over gloo in the CPU tests).
Philox streams are keyed by the GLOBAL env index (DmpState.env_base), so results do not depend on the
number of shards."""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def shard_bounds(total_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """(env_base, count) of rank's contiguous slice; the remainder goes to the lowest ranks."""
    if not (0 <= rank < world) or total_envs < 0:
        raise ValueError("bad shard request")
    q, r = divmod(total_envs, world)
    count = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, count


def allreduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of the float64[4] statistics vector over the process group (no-op when
    torch.distributed is not initialised or world_size == 1)."""
    import torch.distributed as dist
    if stats.dtype != torch.float64 or stats.numel() != 4:
        raise ValueError("stats must be a float64 tensor of 4 elements")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def make_sharded_env(dim: int, total_envs: int, rank: Optional[int] = None, world: Optional[int] = None,
                     device=None, **kw):
    """BatchedDMPEnv over this rank's slice of `total_envs` (rank/world default to torch.distributed's)."""
    import torch.distributed as dist
    from .vecenv import BatchedDMPEnv
    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
        else:
            rank, world = 0, 1
    base, count = shard_bounds(total_envs, rank, world)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return BatchedDMPEnv(dim, num_envs=count, env_base=base, device=device, **kw)
