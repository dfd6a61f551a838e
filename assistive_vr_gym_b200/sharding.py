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


def reduce_max(t, group=None) -> float:
    """Max of a 1-element tensor over ranks: bench.py's "time = max over ranks"."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def make_shard(env_id: str, n_total: int, rank: int, world: int, device: int = 0, seed: int = 1001, **kw):
    """This rank's shard of a global batch of `n_total` environments of `env_id`: a BatchedAssistiveEnv over the contiguous
    range shard_range(n_total, rank, world), seeded per rank.  Shards never exchange data on the step path (one Bullet world per
    environment in the reference, env.py:23); `reduce_episode_stats` is the only collective of a rollout."""
    from .envs import BatchedAssistiveEnv
    begin, end = shard_range(n_total, rank, world)
    env = BatchedAssistiveEnv(env_id, num_envs=end - begin, device=device, seed=seed + rank, **kw)
    env.shard = (begin, end)
    return env
