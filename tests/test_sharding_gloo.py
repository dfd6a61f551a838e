"""N>1 host logic on CPU: env-range sharding and the episode-statistics reduction with world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from assistive_vr_gym_b200.sharding import shard_range, reduce_episode_stats


def test_shard_ranges_partition_the_batch():
    for n, w in [(8, 2), (4096, 8), (10, 4), (7, 8), (32768, 8)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from assistive_vr_gym_b200.envs import load_env_data
    from assistive_vr_gym_b200.compiler.reset import sample_states
    _, resets = load_env_data("ScratchItchJaco.npz")
    n_total = 10
    b, e = shard_range(n_total, rank, world)
    # every rank samples only its shard, seeded by rank like bench.py
    env, var = sample_states(resets, e - b, np.random.RandomState(1001 + rank))
    stats = torch.tensor([float(env[:, 0].sum()), float(var.sum()), 0.0, float(e - b)], dtype=torch.float64)
    local = stats.clone()
    reduce_episode_stats(stats)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    ok = torch.allclose(stats, sum(gathered)) and stats[3].item() == n_total
    # time reduction used by bench.py: max over ranks
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and t.item() == 10.0 + world - 1
    if rank == 0:
        out.put(bool(ok))
    dist.destroy_process_group()


def test_world_size_two_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
