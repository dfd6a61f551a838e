"""Development check of the Feeding / Drinking path on a GPU box: CUDA step vs the CPU oracle from identical states.

    python tools/gpu_fd_check.py [env_id] [n_envs] [n_steps]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import make                                    # noqa: E402
from oracle.oracle import Oracle, env_to_f64, part_to_f64, part_masks      # noqa: E402

env_id = sys.argv[1] if len(sys.argv) > 1 else "FeedingJaco-v0"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
env = make(env_id, num_envs=n, device=0, seed=7)
t = time.time()
obs0 = env.reset().cpu().numpy().copy()
torch.cuda.synchronize()
print(f"{env_id}: reset of {n} envs (device draws + IK + 100 settle steps) {time.time() - t:.2f} s")
state = env.get_state(); part = env.get_particles()
npart = env.sim.n_particles
print("particles z (env 0):", np.round(part[0, 128:128 + min(npart, 8)], 4), "contacts", part[0].view(np.int32)[590], "overflow", part[0].view(np.int32)[591])
print("IK error (env 0..3):", state[:4, 124 + 24])
oracles = [Oracle(b) for b in env.blobs]
recs = [env_to_f64(state[e]).copy() for e in range(n)]
parts = [part_to_f64(part[e]).copy() for e in range(n)]
worst_obs0 = max(float(np.abs(oracles[int(env.variants[e])].reset_obs(recs[e]) - obs0[e]).max()) for e in range(n))
print("reset obs: max |cuda - oracle| =", worst_obs0)
rng = np.random.RandomState(0)
env.sim.enable_debug(True)
for s in range(steps):
    a = rng.uniform(-1, 1, (n, env.sim.n_actions)).astype(np.float32)
    obs, rew, done, info = env.step(torch.as_tensor(a, device="cuda:0"))
    obs = obs.cpu().numpy(); rew = rew.cpu().numpy()
    gstate = env.get_state(); gpart = env.get_particles()
    terms = env.sim.get_reward_terms()
    dq = dx = dr = 0.0
    ev_ok = True
    for e in range(n):
        o = oracles[int(env.variants[e])]
        oobs, orew, oinfo, oc = o.step(recs[e], a[e], parts[e])
        gp = part_to_f64(gpart[e])
        dq = max(dq, float(np.abs(env_to_f64(gstate[e])[:32] - recs[e][:32]).max()))
        alive = part_masks(parts[e], 576)
        idx = [p for p in range(npart) if (alive >> p) & 1]
        if idx:
            dx = max(dx, max(float(np.abs(gp[64 * c + p] - parts[e][64 * c + p])) for c in range(3) for p in idx))
        dr = max(dr, abs(orew - float(rew[e])) if abs(terms[e, 6] - oinfo[6]) < 1e-3 else 0.0)
        for slot in (576, 578, 584, 586, 588):
            if part_masks(gp, slot) != part_masks(parts[e], slot):
                ev_ok = False
                print(f"  step {s} env {e}: mask slot {slot} differs: cuda {part_masks(gp, slot):x} oracle {part_masks(parts[e], slot):x}")
    print(f"step {s}: max |dq| {dq:.2e}  max |dx_particle| {dx:.2e}  max |dreward| {dr:.2e}  events equal {ev_ok}  "
          f"pcontacts cuda {gpart[0].view(np.int32)[590]} oracle {int(parts[0][590])}  reward[0] {rew[0]:.4f}")
print("launches", env.sim.launch_count)
# throughput
for nb in (4096,):
    benv = make(env_id, num_envs=nb, device=0, seed=1)
    t = time.time(); benv.reset(); torch.cuda.synchronize(); t_reset = time.time() - t
    acts = torch.rand((nb, benv.sim.n_actions), device="cuda:0") * 2 - 1
    for _ in range(3):
        benv.step(acts)
    torch.cuda.synchronize()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    K = 20
    for _ in range(K):
        benv.step(acts)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / K
    print(f"{env_id} {nb} envs: {ms:.3f} ms/step = {nb / ms * 1e3:.3e} env-steps/s (reset {t_reset:.2f} s); overflow envs {int((benv.contact_overflow() != 0).sum())}")
