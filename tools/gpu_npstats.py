"""narrowphase counters over windows of an episode (AVG_DBG=32)."""
import os, sys
os.environ["AVG_DBG"] = "32"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import make
E = 32768
env = make("ScratchItchJaco-v0", num_envs=E, device=0, seed=1001)
env.reset()
gen = torch.Generator(device="cuda"); gen.manual_seed(0)
for k in range(150): env.step(torch.rand((E, 7), device="cuda", generator=gen) * 2 - 1)
torch.cuda.synchronize()
st = env.get_state()
env.close()
# second phase: continue from the late state with fresh counters
env = make("ScratchItchJaco-v0", num_envs=E, device=0, seed=1001)
env.reset()
env.sim.set_state(st, env.variants)
for k in range(10): env.step(torch.rand((E, 7), device="cuda", generator=gen) * 2 - 1)
torch.cuda.synchronize()
print(f"10 steps from the state after 150 steps ({E*10*5} env-substeps):", flush=True)
env.close()
