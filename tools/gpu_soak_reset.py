"""Soak of the device reset paths: staggered episodes (auto_reset="device"), on-device IK, policy actions, many episodes.
usage: python tools/gpu_soak_reset.py [n_env=131072] [steps=1000] [ids...]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import make
from assistive_vr_gym_b200.policy import synthetic_policy
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ids = sys.argv[3:] or ["ScratchItchJaco-v0", "ScratchItchPR2-v0", "BedBathingPR2-v0"]
for env_id in ids:
    env = make(env_id, num_envs=n, device=0, seed=9, auto_reset="device", device_ik=True)
    blob, _ = synthetic_policy(env.obs_robot_len, env.action_robot_len, seed=2); env.set_policy(blob)
    env.reset_device(seed=77)
    # stagger: restart a random quarter of the batch at steps 40, 80, 120
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    ends = torch.zeros(n, dtype=torch.int32, device="cuda"); ret = torch.zeros(n, device="cuda"); rets = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(1, T + 1):
        a = env.act() if (t // 100) % 2 == 0 else torch.rand((n, env.sim.n_actions), device="cuda", generator=g) * 2 - 1
        obs, rew, done, info = env.step(a)
        ends += done.to(torch.int32)
        if t in (40, 80, 120):
            env.reset_device(mask=(torch.rand(n, device="cuda", generator=g) < 0.25))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = env.get_state()
    bad = int((~np.isfinite(st[:, :64])).any(axis=1).sum())
    it = st.view(np.int32)[:, 152]
    ikerr = st[:, 127]
    print(f"{env_id}: {n} envs x {T} steps staggered, {n * T / dt:.3e} env-steps/s incl. resets; episodes ended per env min/mean/max "
          f"{int(ends.min())}/{float(ends.float().mean()):.2f}/{int(ends.max())}; iteration counters in [{it.min()}, {it.max()}]; non-finite envs {bad}; "
          f"obs finite {bool(torch.isfinite(obs).all())}; IK error of the current episodes: max {ikerr.max():.3e}, within 0.03: {(ikerr < 0.03).mean() * 100:.2f} %", flush=True)
    env.close()
