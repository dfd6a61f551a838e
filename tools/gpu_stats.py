"""Per-env diagnostics after a few steps: solver iterations, narrowphase candidates, contacts."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import make
n = 16384
env = make("ScratchItchJaco-v0", num_envs=n, device=0, seed=1001)
env.sim.enable_debug(True)
env.reset()
gen = torch.Generator(device="cuda"); gen.manual_seed(0)
for t in range(30):
    env.step(torch.rand((n, 7), device="cuda", generator=gen) * 2 - 1)
    if t in (0, 4, 29):
        st = env.get_state().view(np.int32)
        c, nc = env.sim.get_contacts()
        it = st[:, 167] / 5.0; cand = st[:, 168] / 5.0
        print(f"step {t}: solver iters/substep mean {it.mean():.1f} p50 {np.median(it):.0f} p90 {np.percentile(it,90):.0f} max {it.max():.0f} | "
              f"candidates/substep mean {cand.mean():.1f} p50 {np.median(cand):.0f} p90 {np.percentile(cand,90):.0f} max {cand.max():.0f} | "
              f"contacts mean {nc.mean():.2f} p90 {np.percentile(nc,90):.0f} max {nc.max()} | overflow {np.bitwise_or.reduce(st[:,166])}")
