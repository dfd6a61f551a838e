"""Short Feeding / Drinking run for ncu: reset (200 internal settle steps) + a few env-steps at a given batch size."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import make
env_id = sys.argv[1]; n = int(sys.argv[2]); steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
env = make(env_id, num_envs=n, device=0, seed=1001)
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
a = torch.empty((n, env.sim.n_actions), device="cuda")
for k in range(steps):
    a.uniform_(-1, 1, generator=g); env.step(a)
torch.cuda.synchronize()
print("ok", env.sim.launch_count)
