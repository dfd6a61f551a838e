#!/usr/bin/env python
"""bench.py — env-steps/sec of the ScratchItchJaco-v0 hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs-per-gpu E] [--impl reference]

A "step" is one env.step() of every environment of the batch (take_step + 5 physics sub-steps + reward/obs), with
synthetic uniform random actions (seed 0).  The batch is at STAGGERED EPISODE PHASES: before the timed region it is rolled
forward (untimed) with one tenth of the environments restarted every 20 steps, so that the K timed steps see episode steps
0..199 uniformly -- the mean of whole 200-step random-action episodes, which is also what the reference arm runs (steps right
after a reset have fewer contacts and would flatter the simulator; that window is reported as `post_reset_window`).
`value` is measured with the
state, actions and outputs resident in HBM (CUDA events on the launch stream, max over ranks); `e2e` is the same
metric through the reference-facing host-buffer call (avg_step_host: NumPy actions in, NumPy obs/reward/done/info out,
host<->device copies inside the timed region).  Environments are sharded over ranks with no data-path collective
("weak" scaling, fixed envs per GPU); NCCL is used only for the barrier, the max-time reduction and the episode
statistics.  `--impl reference` times the CPU oracle port of the path (the reference's own arithmetic lives in the
pybullet extension, which is not installable here) on all host cores for the same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENV_ID = "ScratchItchJaco-v0"
METRIC = "env-steps/sec (ScratchItchJaco, 1/2/4/8 B200) vs PyBullet on host cores"
UNIT = "env-steps/s"


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_hash() -> str:
    """sha256/16 of the kernel sources: profiles/traffic.json records the hash it was captured at; a stale capture is not
    reported as this build's traffic."""
    import hashlib
    h = hashlib.sha256()
    for f in ("assistive_vr_gym_b200/csrc/avg_kernels.cu", "assistive_vr_gym_b200/csrc/avg_kernels.h", "assistive_vr_gym_b200/csrc/avg_math.cuh",
              "assistive_vr_gym_b200/csrc/avg_capi.cu", "include/avg_model.h"):
        with open(os.path.join(ROOT, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
            except Exception:
                continue
            for n, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _oracle_worker(args):
    blob, recs, actions = args
    from oracle.oracle import Oracle
    o = Oracle(blob)
    t0 = time.perf_counter()
    n = 0
    for e in range(len(recs)):
        rec = recs[e].copy()
        for t in range(actions.shape[0]):
            o.step(rec, actions[t, e])
            n += 1
    return n, time.perf_counter() - t0


def cpu_oracle_throughput(envs_per_core: int, steps: int, cores: int | None = None):
    """Times the CPU oracle (kind 'port') on `cores` processes; returns (env-steps/s, cores, sample text)."""
    import multiprocessing as mp
    from assistive_vr_gym_b200.envs import load_env_data
    from assistive_vr_gym_b200.compiler.reset import sample_states
    from oracle.oracle import env_to_f64, build
    build()
    cores = cores or os.cpu_count() or 1
    blobs, resets = load_env_data("ScratchItchJaco.npz")
    rng = np.random.RandomState(1001)
    jobs = []
    for c in range(cores):
        env, variant = sample_states(resets, envs_per_core, rng, genders=np.full(envs_per_core, c % len(blobs)))
        acts = np.random.RandomState(c).uniform(-1, 1, (steps, envs_per_core, 7)).astype(np.float32)
        jobs.append((blobs[c % len(blobs)], env_to_f64(env), acts))
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_oracle_worker, jobs)
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return total / busy, cores, f"{cores} processes x {envs_per_core} envs x {steps} steps of {ENV_ID} (random actions), {wall:.1f} s wall"


def run_reference(args, rank: int, world: int):
    """The reference arm: the CPU restatement of the path (kind "port": PyBullet is not installable here) on every host core.
    One bench "step" = one bounded sample: 16 whole 200-step random-action episodes per core (about a second), so that the
    K timed repetitions it reports as `steps` are the ones it ran and K x ms_per_step is the run's wall time."""
    if rank != 0:
        return
    import ctypes
    from oracle.oracle import build as build_oracle
    ctypes.CDLL(build_oracle())                  # mapped in this process too (the workers are forked from it)
    cores = os.cpu_count() or 1
    per_core, ep_steps = 16, 200
    for _ in range(max(args.warmup, 1)):
        cpu_oracle_throughput(1, 20, cores)
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, c, sample = cpu_oracle_throughput(per_core, ep_steps, cores)
        vals.append(v)
    ms_per = (time.perf_counter() - t0) * 1e3 / max(args.steps, 1)
    value = float(np.median(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{ENV_ID}, whole 200-step random-action episodes, CPU oracle port of the path (PyBullet itself is not installable here)",
                       "step": f"one step = {cores} processes x {per_core} episodes x {ep_steps} env-steps"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample + f", median of {args.steps}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


_REAL_STDOUT = None


def _capture_stdout():
    """Everything libraries write to fd 1 from here on (NCCL prints its version banner there) goes to stderr; the JSON line
    is written to the saved descriptor, so stdout carries exactly ONE line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--envs-per-gpu", type=int, default=393216)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-episode", action="store_true", help="skip the extra whole-episode (200-step) throughput measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from assistive_vr_gym_b200 import make
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    from assistive_vr_gym_b200.sharding import make_shard, reduce_episode_stats, reduce_max
    E = args.envs_per_gpu
    # weak scaling: every rank owns E environments of the global batch [rank * E, (rank + 1) * E) (sharding.make_shard)
    env = make_shard(ENV_ID, E * world, rank, world, device=local_rank, seed=1001)
    assert env.num_envs == E
    env.reset()
    gen = torch.Generator(device=dev); gen.manual_seed(rank)
    # fresh i.i.d. actions for every step (reference examples/random_actions.py samples anew each step; a short ring
    # reused cyclically would give every arm a constant drift and pile the batch up against limits and obstacles)
    n_ring = max(args.warmup, 3) + args.steps
    ring = [torch.rand((E, 7), device=dev, generator=gen) * 2 - 1 for _ in range(min(n_ring, 256))]
    act_ep = torch.empty((E, 7), device=dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    W = max(args.warmup, 3)
    # ---- post-reset window (extra): the K steps right after a reset, few contacts yet -----------------------------
    for w in range(W):
        env.step(ring[w % len(ring)])
    env.elapsed = 0
    barrier()
    q0 = torch.cuda.Event(enable_timing=True); q1 = torch.cuda.Event(enable_timing=True)
    q0.record(stream)
    for k in range(args.steps):
        env.step(ring[(W + k) % len(ring)]); env.elapsed = 0
    q1.record(stream)
    barrier()
    post_reset_ms = reduce_max(torch.tensor([q0.elapsed_time(q1)], device=dev))
    # ---- staggered episode phases: group g (a tenth of the batch) restarts at step 20 g of a 200-step untimed roll, so the
    #      timed window holds environments at episode steps 0..199 in equal shares (the whole-episode mean) --------------
    env.reset()
    groups = 10
    gid = torch.arange(E, device=dev) % groups
    for k in range(200 - W):
        if k % 20 == 0 and k > 0:
            env.reset_device(mask=(gid == (k // 20)))
        act_ep.uniform_(-1, 1, generator=gen)
        env.step(act_ep); env.elapsed = 0
    # the state the timed window starts from (+ W warm-up steps): the end-to-end measurement below restarts from it, so both
    # numbers cover the SAME steps of the same episodes (a batch that just rolls on gets slower: more environments in contact)
    snap_state = env.get_state(); snap_variants = env.sim.get_variants().copy()
    for w in range(W):
        env.step(ring[w % len(ring)]); env.elapsed = 0
    barrier()
    sampler = ClockSampler(local_rank); sampler.start()
    launches0 = env.sim.launch_count
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    stats = torch.zeros(4, device=dev)
    for k in range(args.steps):
        obs, rew, done, info = env.step(ring[(W + k) % len(ring)])
        env.elapsed = 0                              # keep the timed window free of TimeLimit resets
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = env.sim.launch_count - launches0
    sampler.stop_flag = True; sampler.join(timeout=2)
    # episode statistics: the only collective of the path (NCCL all-reduce of a 4-float vector)
    stats[0] = rew.sum(); stats[1] = info["task_success"].sum(); stats[2] = info["total_force_on_human"].sum(); stats[3] = float(E)
    overflow_envs = float((info["contact_overflow"] != 0).sum())
    ms = reduce_max(torch.tensor([ms], device=dev))
    reduce_episode_stats(stats)
    value = E * world * args.steps / (ms * 1e-3)

    # ---- end-to-end through the host-buffer API: the same steps of the same episodes as the device-resident measurement
    #      (the staggered batch restored from its snapshot, the same W warm-up steps and actions, then the timed steps), so the
    #      two numbers differ only by the host<->device traffic and the host-side call.  Actions sit in page-locked arrays (a caller writes its policy output there); results
    #      come back in the env's own pinned arrays.
    from assistive_vr_gym_b200 import capi
    e2e_steps = max(3, min(args.steps, 20))
    a_pin = [capi.PinnedArray((E, 7), np.float32) for _ in range(min(W + e2e_steps, 32))]
    for i, p in enumerate(a_pin):
        p.array[...] = ring[i % len(ring)].cpu().numpy()
    env.set_state(snap_state, snap_variants)         # the staggered batch as it was before the device-resident measurement
    del snap_state
    for w in range(W):
        env.step_host(a_pin[w % len(a_pin)].array); env.elapsed = 0
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        o, r, d, i = env.step_host(a_pin[(W + k) % len(a_pin)].array); env.elapsed = 0
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    e2e_value = E * world * e2e_steps / (reduce_max(torch.tensor([e2e_ms], device=dev)) * 1e-3)
    h2d = E * 7 * 4
    d2h = E * (env.sim.n_obs * 4 + 4 + 8 + 1)

    # ---- whole-episode throughput (extra, not the headline): random-action episodes get slower as arms wander into
    #      contact (more candidate pairs, PGS sweeps up to the 50-iteration cap), so the steps right after reset flatter
    #      the simulator; this is the mean over a full TimeLimit(200) episode from a fresh reset -----------------------
    episode = None
    if not args.no_episode:
        env.reset()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(200):
            act_ep.uniform_(-1, 1, generator=gen)            # one tiny RNG kernel per step inside the timed region
            env.step(act_ep); env.elapsed = 0
        e1.record(stream)
        barrier()
        tep = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if distributed:
            dist.all_reduce(tep, op=dist.ReduceOp.MAX)
        episode = {"value": E * world * 200 / (float(tep.item()) * 1e-3), "unit": UNIT, "steps": 200,
                   "note": "mean over a full 200-step random-action episode from reset (device-resident I/O)"}

    # ---- synthetic-policy rollout (extra): obs -> MLP actor -> action -> step, all on the device, from a fresh reset
    #      (north_star asks for random-action AND policy rollouts; no checkpoint ships with the reference) ---
    policy = None
    if not args.no_episode:
        from assistive_vr_gym_b200.policy import synthetic_policy
        blob, _ = synthetic_policy(env.obs_robot_len, env.action_robot_len, seed=0)
        env.set_policy(blob)
        env.reset()
        env.rollout(3); env.elapsed = 0
        barrier()
        p0 = torch.cuda.Event(enable_timing=True); p1 = torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for k in range(50):
            env.step(env.act()); env.elapsed = 0
        p1.record(stream)
        barrier()
        tpo = torch.tensor([p0.elapsed_time(p1)], device=dev)
        if distributed:
            dist.all_reduce(tpo, op=dist.ReduceOp.MAX)
        policy = {"value": E * world * 50 / (float(tpo.item()) * 1e-3), "unit": UNIT, "steps": 50,
                  "note": "synthetic-policy rollout: orthogonal-init 30-64-64-7 tanh actor on VecNormalize'd observations, inference fused on "
                          "the device (avg_policy_act), steps 3-52 after reset"}

    # ---- batch-size sweep (extra): the same random-action steps at smaller batches; 4096 envs is less than one wave of
    #      warps on 148 SMs, so small batches are latency-bound by the 44-kernel launch sequence ------------------------------
    sweep = None
    if not args.no_episode:
        sweep = []
        for nb, graphed in ((1024, False), (1024, True), (4096, False), (4096, True), (16384, False), (16384, True),
                            (65536, False), (196608, False)):
            # cuda_graph=True: the step's launch sequence replayed as one CUDA graph (envs.py) -- where launches bind
            senv_b = make(ENV_ID, num_envs=nb, device=local_rank, seed=1001 + rank, cuda_graph=graphed)
            senv_b.reset_device(seed=1001 + rank)
            sa = torch.empty((nb, 7), device=dev)
            nsw = 20 if nb > 16384 else 100
            for k in range(3):
                sa.uniform_(-1, 1, generator=gen); senv_b.step(sa); senv_b.elapsed = 0
            barrier()
            s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
            s0.record(stream)
            for k in range(nsw):
                sa.uniform_(-1, 1, generator=gen); senv_b.step(sa); senv_b.elapsed = 0
            s1.record(stream)
            barrier()
            ts = torch.tensor([s0.elapsed_time(s1)], device=dev)
            if distributed:
                dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            sweep.append({"envs_per_gpu": nb, "cuda_graph": graphed, "value": nb * world * nsw / (float(ts.item()) * 1e-3), "unit": UNIT, "steps": nsw})
            senv_b.close()

    # ---- BedBathing (extra; BASELINE.json configs[2] names BedBathingPR2-v0 with a pretrained policy at 8192 envs: no
    #      checkpoint ships, so the policy is the synthetic one): whole 200-step episodes from a device reset, on the PR2
    #      at the named batch size, on the Jaco at the named and at a GPU-filling batch size ------------------------------
    bed = None
    if not args.no_episode:
        from assistive_vr_gym_b200.policy import synthetic_policy
        bed = []
        for bid, nb in (("BedBathingPR2-v0", 8192), ("BedBathingPR2-v0", 65536), ("BedBathingJaco-v0", 8192), ("BedBathingJaco-v0", 65536),
                        ("ScratchItchPR2-v0", 65536)):
            benv = make(bid, num_envs=nb, device=local_rank, seed=1001 + rank, cuda_graph=nb <= 16384)   # graph replay where launches matter
            pblob, _ = synthetic_policy(benv.obs_robot_len, benv.action_robot_len, seed=0)
            benv.set_policy(pblob)
            benv.reset_device(seed=1001 + rank)
            benv.rollout(3); benv.elapsed = 0
            benv.reset_device(seed=1001 + rank)
            barrier()
            b0 = torch.cuda.Event(enable_timing=True); b1 = torch.cuda.Event(enable_timing=True)
            b0.record(stream)
            for k in range(200):
                bo, br, bd, bi = benv.step(benv.act()); benv.elapsed = 0
            b1.record(stream)
            barrier()
            tb = torch.tensor([b0.elapsed_time(b1)], device=dev)
            bstat = torch.stack([bi["task_success"].float().sum(), benv.reward.sum()])
            if distributed:
                dist.all_reduce(tb, op=dist.ReduceOp.MAX)
                dist.all_reduce(bstat, op=dist.ReduceOp.SUM)
            bed.append({"env_id": bid, "envs_per_gpu": nb, "cuda_graph": benv.cuda_graph, "value": nb * world * 200 / (float(tb.item()) * 1e-3), "unit": UNIT,
                        "steps": 200, "task_success_rate": float(bstat[0]) / (nb * world), "mean_reward_last_step": float(bstat[1]) / (nb * world),
                        "note": "synthetic-policy rollout (%d-64-64-7 tanh actor, fused inference), full episode from a device reset" % benv.obs_robot_len})
            benv.close()

    # ---- BASELINE.json configs[1] (FeedingSawyer-v0, 4096 envs on one B200, random actions) and configs[3] (DrinkingBaxter-v0,
    #      4096 envs per GPU, env-sharded: under torchrun every rank owns 4096): whole episodes (Feeding) / the first 50 steps
    #      (Drinking: 64 water particles, ~250 particle contacts per internal step) from a device reset.  Both ids are
    #      build-defined (the reference's task files have no Sawyer / Baxter branch) ----------------------------
    if not args.no_episode:
        for fid, nb, nst in (("FeedingSawyer-v0", 4096, 200), ("FeedingJaco-v0", 4096, 200), ("DrinkingBaxter-v0", 4096, 50), ("DrinkingJaco-v0", 4096, 50)):
            fenv = make_shard(fid, nb * world, rank, world, device=local_rank, seed=1001, cuda_graph=True)   # 52 launches per env-step: one graph replay
            fenv.reset()
            fa = torch.empty((nb, fenv.sim.n_actions), device=dev)
            for k in range(3):
                fa.uniform_(-1, 1, generator=gen); fenv.step(fa); fenv.elapsed = 0
            fenv.reset()
            barrier()
            f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
            f0.record(stream)
            for k in range(nst):
                fa.uniform_(-1, 1, generator=gen); fo, fr, fd_, fi = fenv.step(fa); fenv.elapsed = 0
            f1.record(stream)
            barrier()
            tf_ = reduce_max(torch.tensor([f0.elapsed_time(f1)], device=dev))
            fstat = torch.stack([fi["task_success"].float().sum(), fr.sum(), (fi["contact_overflow"] != 0).float().sum(), torch.tensor(float(nb), device=dev)])
            reduce_episode_stats(fstat)
            bed.append({"env_id": fid, "envs_per_gpu": nb, "value": nb * world * nst / (tf_ * 1e-3), "unit": UNIT, "steps": nst,
                        "task_success_rate": float(fstat[0] / fstat[3]), "mean_reward_last_step": float(fstat[1] / fstat[3]),
                        "envs_with_contact_overflow": int(fstat[2]),
                        "cuda_graph": True,
                        "note": "random actions, 5 frames x 2 internal steps x 10 PGS iterations per env-step, %d particles, from a device reset "
                                "(fresh IK start pose + 100 settle steps per episode)" % fenv.sim.n_particles})
            fenv.close()

    # ---- a `New` id (extra): ScratchItchJacoNew-v0, random-action episode from a device reset (32 persons of different height /
    #      waist pose as model variants, the arm pose drawn per episode with the collision-free resampling on the device) ----------
    if not args.no_episode:
        nenv = make("ScratchItchJacoNew-v0", num_envs=65536, device=local_rank, seed=1001 + rank)
        nenv.reset_device(seed=1001 + rank)
        na_ = torch.empty((65536, 7), device=dev)
        for k in range(3):
            na_.uniform_(-1, 1, generator=gen); nenv.step(na_); nenv.elapsed = 0
        nenv.reset_device(seed=1001 + rank)
        barrier()
        n0 = torch.cuda.Event(enable_timing=True); n1 = torch.cuda.Event(enable_timing=True)
        n0.record(stream)
        for k in range(200):
            na_.uniform_(-1, 1, generator=gen); no_, nr_, nd_, ni_ = nenv.step(na_); nenv.elapsed = 0
        n1.record(stream)
        barrier()
        tn = reduce_max(torch.tensor([n0.elapsed_time(n1)], device=dev))
        bed.append({"env_id": "ScratchItchJacoNew-v0", "envs_per_gpu": 65536, "value": 65536 * world * 200 / (tn * 1e-3), "unit": UNIT, "steps": 200,
                    "envs_with_contact_overflow": int((ni_["contact_overflow"] != 0).sum()),
                    "note": "random actions, full episode from a device reset; 32 model variants (persons) per batch"})
        nenv.close()

    # ---- mixed-task batch (extra; BASELINE.json configs[3] "mixed-task batch (all robots)"): one homogeneous sub-batch per
    #      (task, robot) id, each its own handle, stepped concurrently on separate streams (assistive_vr_gym_b200/mixed.py) ----
    if not args.no_episode:
        from assistive_vr_gym_b200.mixed import MixedBatch
        per_id = 2048
        mixed_ids = [t + r + "-v0" for t, robots in (("ScratchItch", ("Jaco", "PR2")), ("BedBathing", ("Jaco", "PR2")),
                                                     ("Feeding", ("Jaco", "PR2", "Sawyer", "Baxter")), ("Drinking", ("Jaco", "PR2", "Sawyer", "Baxter")))
                     for r in robots]
        mb = MixedBatch(env_ids=mixed_ids, envs_per_id=per_id, device=local_rank, seed=1001 + 100 * rank)
        mb.reset_device(seed=1001 + rank)
        for k in range(3):
            mb.step(mb.sample_actions(gen))
            for e_ in mb.envs.values():
                e_.elapsed = 0
        barrier()
        m0 = torch.cuda.Event(enable_timing=True); m1 = torch.cuda.Event(enable_timing=True)
        m0.record(stream)
        for k in range(50):
            mb.step(mb.sample_actions(gen))
            for e_ in mb.envs.values():
                e_.elapsed = 0
        m1.record(stream)
        barrier()
        tm = torch.tensor([reduce_max(torch.tensor([m0.elapsed_time(m1)], device=dev))])
        bed.append({"env_id": "mixed: " + ", ".join(mb.env_ids), "envs_per_gpu": mb.num_envs, "value": mb.num_envs * world * 50 / (float(tm.item()) * 1e-3),
                    "unit": UNIT, "steps": 50, "note": "%d ids x %d envs, one handle and one stream per id, random actions, steps 3-52 after a device reset" % (len(mb.env_ids), per_id)})
        mb.close()

    # ---- BASELINE.json configs[4] (ScratchItchJacoHuman-v0, 4096 envs per GPU, both halves of the action driven) and
    #      configs[0] (one ScratchItchJaco-v0 environment through the reference-typed NumPy API, 200 steps) as extras --------
    if not args.no_episode:
        henv = make("ScratchItchJacoHuman-v0", num_envs=4096, device=local_rank, seed=1001 + rank)
        henv.reset_device(seed=1001 + rank)
        ha = torch.empty((4096, 17), device=dev)
        for k in range(3):
            ha.uniform_(-1, 1, generator=gen); henv.step(ha); henv.elapsed = 0
        henv.reset_device(seed=1001 + rank)
        barrier()
        h0 = torch.cuda.Event(enable_timing=True); h1 = torch.cuda.Event(enable_timing=True)
        h0.record(stream)
        for k in range(200):
            ha.uniform_(-1, 1, generator=gen); henv.step(ha); henv.elapsed = 0
        h1.record(stream)
        barrier()
        th = torch.tensor([h0.elapsed_time(h1)], device=dev)
        if distributed:
            dist.all_reduce(th, op=dist.ReduceOp.MAX)
        bed.append({"env_id": "ScratchItchJacoHuman-v0", "envs_per_gpu": 4096, "value": 4096 * world * 200 / (float(th.item()) * 1e-3), "unit": UNIT,
                    "steps": 200, "note": "random robot + human actions (17), arm-limit MLP after every sub-step, full episode from a device reset"})
        henv.close()
        if rank == 0:
            senv = make("ScratchItchJaco-v0", device=local_rank, seed=1001)       # num_envs=None: float64 obs, python scalars, like gym.make
            senv.reset()
            sact = np.random.RandomState(0).uniform(-1, 1, (200, 7)).astype(np.float32)
            for k in range(5):
                senv.step(sact[k])
            senv.reset()
            t0 = time.perf_counter()
            for k in range(200):
                senv.step(sact[k])
            bed.append({"env_id": "ScratchItchJaco-v0", "envs_per_gpu": 1, "value": 200 / (time.perf_counter() - t0), "unit": UNIT, "steps": 200,
                        "note": "ONE environment through the reference-typed NumPy API (examples/random_actions.py loop): launch-latency bound"})
            senv.close()

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        bpe = env.sim.bytes_per_env_step
        # one "launch" of the path = the kernel sequence of one env-step (prologue, 5 x {collide, narrowphase, dynamics,
        # solve}, epilogue); its algorithmic bytes are per env-step, its duration the step time
        achieved = bpe * E / (ms / args.steps * 1e-3) / 1e9
        traffic = None; prof = {}; traffic_note = "no committed ncu capture"
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                prof = json.load(open(tp))
                if prof.get("kernel_hash") != kernel_source_hash():
                    # a capture of other kernels says nothing about this build: refuse it
                    traffic_note = "profiles/traffic.json was captured at kernel hash %s, this build is %s: not reported" % (prof.get("kernel_hash"), kernel_source_hash())
                    prof = {}
                else:
                    traffic = prof.get("dram_bytes_per_step_per_env", None)
                    traffic = traffic * E if traffic is not None else None
                    traffic_note = "ncu dram__bytes of the step's kernels (%s, %d envs) x this batch" % (prof.get("source"), prof.get("n_env", 0))
            except Exception:
                traffic = None; prof = {}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{ENV_ID}, {E} envs/GPU, uniform random actions (seed 0), staggered episode phases (a tenth of the batch "
                                       f"restarted every 20 steps: the timed steps cover episode steps 0..199 uniformly = whole-episode mean), "
                                       f"5 sub-steps x 50 PGS iterations per env-step",
                           "reset": "device sampler: every episode draws its own start target and solves the IK on the GPU (util.py:34-57 incl. the "
                                    "5-step self-contact test)",
                           "envs_per_gpu": E, "l2": "state + I/O per GPU = %.0f MB > 126 MB L2 (inputs larger than L2)" % ((E * (768 + 37 * 4 + 13)) / 1e6),
                           "parallelism": f"env-sharded x{world}, no step-path collective"},
                "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps},
                "post_reset_window": {"value": E * world * args.steps / (post_reset_ms * 1e-3), "unit": UNIT, "steps": args.steps,
                                      "note": "the same K steps taken right after a reset (few contacts yet): NOT the headline"},
                "envs_with_contact_overflow": int(overflow_envs),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                             "peak_source": peak_src, "bytes_per_env_step": bpe, "launch": "the %d kernel launches of one env-step (prologue, 5 x (collide, narrowphase, dynamics, solve), epilogue; batches of 2048 .. 131071 environments step as two halves on two streams)" % (launches // max(1, args.steps)),
                             "issue_slots": (lambda s: None if not s else dict(s, achieved_warp_inst_per_s=s["warp_inst_per_env_step"] * value / world,
                                                                                  frac=s["warp_inst_per_env_step"] * value / world / s["peak_warp_inst_per_s"]))(prof.get("issue_slots")),
                             "fp32": (lambda fp: None if not fp else {"flop_per_env_step": fp["flop_per_env_step_substep_kernels"],
                                                                       "achieved_tflops": fp["flop_per_env_step_substep_kernels"] * value / world / 1e12,
                                                                       "peak_tflops": 148 * 128 * 2 * 1.965e-3,
                                                                       "frac": fp["flop_per_env_step_substep_kernels"] * value / world / 1e12 / (148 * 128 * 2 * 1.965e-3),
                                                                       "note": "FADD + FMUL + 2 FFMA thread instructions per env-step from the committed ncu full-set capture "
                                                                               "(profiles/traffic.json) x measured env-steps/s; FMA-pipe utilisation per kernel is in profiles/ncu_full_r2e.txt"})(prof.get("fp32")),
                             "note": "issue/latency bound by design (SURVEY.md 8d): the HBM fraction is reported as north_star asks; "
                                     "issue_slots (from the committed ncu launch list) is the roof that binds"},
                "clocks": sampler.summary(),
                "episode_stats": {"mean_reward_last_step": float(stats[0] / stats[3]), "task_success": float(stats[1]),
                                  "mean_force_on_human": float(stats[2] / stats[3])}}
        if episode is not None:
            line["episode"] = episode
        if policy is not None:
            line["policy_rollout"] = policy
        if sweep is not None:
            line["batch_sweep"] = sweep
        if bed is not None:
            line["other_workloads"] = bed
        if not args.no_cpu_baseline and world == 1:
            v, c, sample = cpu_oracle_throughput(192, 200)         # ~10-20 s of CPU work on the box's cores: whole 200-step episodes
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": c, "kind": "port", "sample": sample}
        _emit(line)
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
