"""ms/step over a whole 200-step episode of random actions (blocks of 10 steps) + diagnostics."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import make
E = int(sys.argv[1]) if len(sys.argv) > 1 else 196608
env = make("ScratchItchJaco-v0", num_envs=E, device=0, seed=1001)
env.sim.enable_debug(True)
env.reset()
gen = torch.Generator(device="cuda"); gen.manual_seed(0)
act = torch.empty((E, 7), device="cuda")
tot = 0.0
for blk in range(20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(10):
        act.uniform_(-1, 1, generator=gen)          # fresh i.i.d. actions every step, like examples/random_actions.py
        env.step(act)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    tot += dt * 10
    st = env.get_state().view(np.int32)
    c, nc = env.sim.get_contacts()
    it = st[:, 167] / 5.0; cand = st[:, 168] / 5.0
    print(f"steps {blk*10:3d}-{blk*10+9:3d}: {dt*1e3:7.2f} ms/step {E/dt/1e6:6.2f} M env-steps/s | iters/substep mean {it.mean():5.1f} p90 {np.percentile(it,90):3.0f} max {it.max():3.0f} | "
          f"cand mean {cand.mean():5.1f} | contacts mean {nc.mean():.2f} p99 {np.percentile(nc,99):.0f} max {nc.max()} | overflow {np.bitwise_or.reduce(st[:,166])}", flush=True)
print(f"whole episode: {E*200/tot/1e6:.2f} M env-steps/s")
