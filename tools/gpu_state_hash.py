"""SHA-1 of the environment records after a seeded random-action roll (compare library variants / switches bit for bit):
python tools/gpu_state_hash.py <env_id> <n_env> <steps>"""
import hashlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import make
env_id, n, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
env = make(env_id, num_envs=n, device=0, seed=1001)
env.reset()
gen = torch.Generator(device="cuda"); gen.manual_seed(0)
act = torch.empty((n, env.sim.n_actions), device="cuda")
tot = 0.0
for k in range(steps):
    act.uniform_(-1, 1, generator=gen); o, r, d, i = env.step(act); env.elapsed = 0
    tot += float(r.sum())
st = env.get_state()
print(f"{env_id} {n} envs {steps} steps: state sha1 {hashlib.sha1(np.ascontiguousarray(st).tobytes()).hexdigest()} reward sum {tot:.6f} "
      f"[AVG_FUSE={os.environ.get('AVG_FUSE', '-')} lib={os.environ.get('AVG_B200_LIB', 'default')}]")
