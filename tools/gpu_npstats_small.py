"""narrowphase counters at a small batch (AVG_DBG=32): usage gpu_npstats_small.py <env_id> <n_env> <warm steps> <steps>"""
import os, sys
os.environ["AVG_DBG"] = "32"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import make
env_id = sys.argv[1]; E = int(sys.argv[2]); warm = int(sys.argv[3]); steps = int(sys.argv[4])
env = make(env_id, num_envs=E, device=0, seed=1001)
env.reset()
gen = torch.Generator(device="cuda"); gen.manual_seed(0)
na = env.sim.n_actions
for k in range(warm): env.step(torch.rand((E, na), device="cuda", generator=gen) * 2 - 1); env.elapsed = 0
torch.cuda.synchronize()
st = env.get_state(); variants = np.asarray(env.variants).copy()
env.close()
env = make(env_id, num_envs=E, device=0, seed=1001)
env.reset()
env.set_state(st, variants)
for k in range(steps): env.step(torch.rand((E, na), device="cuda", generator=gen) * 2 - 1); env.elapsed = 0
torch.cuda.synchronize()
print(f"{steps} steps from the state after {warm} steps ({E*steps*5} env-substeps):", flush=True)
env.close()
