"""Batched use: N environments resident on one GPU, CUDA tensors in and out, episodes restarted on the device with a freshly
solved start pose (on-device IK), actions from a policy evaluated on the device (the loop of the reference's enjoy_vr.py:105-117
without the headset).  usage: python examples/batched_rollout.py [env_id] [num_envs]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import assistive_vr_gym_b200 as assistive_gym
from assistive_vr_gym_b200.policy import synthetic_policy

env_id = sys.argv[1] if len(sys.argv) > 1 else 'ScratchItchJaco-v0'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
env = assistive_gym.make(env_id, num_envs=n, device_ik=True)
blob, _ = synthetic_policy(env.obs_robot_len, env.action_robot_len, seed=0)     # no checkpoint ships with the reference
env.set_policy(blob)
obs = env.reset_device(seed=1)
returns = torch.zeros(n, device=obs.device)
torch.cuda.synchronize(); t0 = time.perf_counter()
for episode in range(2):
    for t in range(200):
        obs, reward, done, info = env.step(env.act())
        returns += reward
    print(f'episode {episode}: mean return {float(returns.mean()):.3f}, task_success {int(info["task_success"].sum())} of {n}, '
          f'mean force on human {float(info["total_force_on_human"].mean()):.3f} N')
    returns.zero_()
    obs = env.reset_device()                         # TimeLimit(200) ends all lock-stepped episodes together; reset_device(mask=...) restarts a subset
torch.cuda.synchronize()
print(f'{2 * 200 * n / (time.perf_counter() - t0):.3e} env-steps/s including resets and policy inference')
env.close()
