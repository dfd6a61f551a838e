"""Soak run: whole episodes at a large batch for every registered id, random actions and the synthetic policy; reports
non-finite states, overflow flags and throughput (development aid: latent data-dependent faults show up at scale).
usage: python tools/gpu_soak.py [n_env=262144] [ids...]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import make
from assistive_vr_gym_b200.envs import REGISTRY
from assistive_vr_gym_b200.policy import synthetic_policy
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
ids = sys.argv[2:] or sorted(REGISTRY)
for env_id in ids:
    particles = REGISTRY[env_id]["task"] in ("feeding", "drinking")
    for mode in (("random",) if particles else ("random", "policy")):
        n_id = min(n, 32768) if particles else n             # 64 water spheres: ~2.7e5 env-steps/s
        env = make(env_id, num_envs=n_id, device=0, seed=5)
        env.reset_device(seed=123 if mode == "random" else 321)
        if mode == "policy":
            blob, _ = synthetic_policy(env.obs_robot_len, env.action_robot_len, seed=1); env.set_policy(blob)
        g = torch.Generator(device="cuda"); g.manual_seed(7)
        a = torch.empty((n_id, env.sim.n_actions), device="cuda")
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for k in range(200):
            if mode == "policy":
                env.step(env.act())
            else:
                a.uniform_(-1, 1, generator=g); env.step(a)
            env.elapsed = 0
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        st = env.get_state()
        bad = int((~np.isfinite(st[:, :64])).any(axis=1).sum())
        ov = np.bincount(st.view(np.int32)[:, 166] & 7, minlength=8)
        print(f"{env_id:26s} {mode:6s} {n_id} envs x 200 steps: {n_id * 200 / dt:.3e} env-steps/s, non-finite envs {bad}, overflow flags (none, contacts>32, rows, both, cand>128...) {ov.tolist()}, "
              f"mean reward {float(env.reward.mean()):.3f}, success {int(env.info_dev[:, 1].sum())}", flush=True)
        env.close()
