"""Per-kernel time split of one env-step (AVG_KERNEL_TIMES=1 development aid of avg_launch_step).
usage: AVG_KERNEL_TIMES=1 python tools/gpu_kernel_times.py <env_id> <n_env> [policy]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import make
from assistive_vr_gym_b200.policy import synthetic_policy
env_id = sys.argv[1]; n = int(sys.argv[2]); use_policy = len(sys.argv) > 3
env = make(env_id, num_envs=n, device=0, seed=1001)
env.reset_device(seed=1001)
if use_policy:
    blob, _ = synthetic_policy(env.obs_robot_len, env.action_robot_len, seed=0); env.set_policy(blob)
g = torch.Generator(device="cuda"); g.manual_seed(0)
a = torch.empty((n, env.sim.n_actions), device="cuda")
for k in range(64):
    if use_policy:
        env.step(env.act())
    else:
        a.uniform_(-1, 1, generator=g); env.step(a)
    env.elapsed = 0
torch.cuda.synchronize()
st = env.get_state().view(np.int32)
print(env_id, n, "mean solver iterations per env-step", st[:, 167].mean(), "mean candidates", st[:, 168].mean())
