"""Environment sharding across GPUs: contiguous env ranges per rank, no step-path traffic; the only
collective is a reduction of a small episode-statistics vector (reference `replay_vr_savemeta.py:46-74` aggregates the
same quantities on the host)."""
from __future__ import annotations

from typing import Tuple

STAT_FIELDS = ("sum_reward", "sum_task_success", "sum_total_force_on_human", "n_envs")


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of the global env index range owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def reduce_episode_stats(stats, group=None):
    """Sum the STAT_FIELDS vector over ranks (NCCL on GPUs, gloo in the CPU tests). Returns the reduced tensor."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats
