"""Bring a batch into a named regime of the episode, then run a few steps between cudaProfilerStart / Stop so that
`ncu --profile-from-start off` captures exactly those steps (the in-contact / whole-episode evidence VERDICT r1 asked for).

usage: python tools/gpu_regime.py <env_id> <n_env> <regime> [steps]
  regime: post_reset | stagger (a tenth of the batch restarted every 20 steps: episode steps 0..199 in equal shares)
          | late (steps 150.. of a random-action episode) | policy (synthetic-policy rollout, steps 40..)
Without ncu it prints the device-timed ms/step of the same steps."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import make
from assistive_vr_gym_b200.policy import synthetic_policy

env_id = sys.argv[1]; n = int(sys.argv[2]); regime = sys.argv[3]; steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
env = make(env_id, num_envs=n, device=0, seed=1001)
env.reset()
gen = torch.Generator(device="cuda"); gen.manual_seed(0)
act = torch.empty((n, env.sim.n_actions), device="cuda")


def rnd_step():
    act.uniform_(-1, 1, generator=gen); env.step(act); env.elapsed = 0


if regime == "post_reset":
    for k in range(3): rnd_step()
elif regime == "stagger":
    gid = torch.arange(n, device="cuda") % 10
    for k in range(200):
        if k % 20 == 0 and k > 0:
            env.reset_device(mask=(gid == (k // 20)))
        rnd_step()
elif regime == "late":
    for k in range(150): rnd_step()
elif regime == "policy":
    blob, _ = synthetic_policy(env.obs_robot_len, env.action_robot_len, seed=0); env.set_policy(blob)
    for k in range(40):
        env.step(env.act()); env.elapsed = 0
else:
    raise SystemExit("unknown regime " + regime)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
for k in range(steps):
    if regime == "policy":
        env.step(env.act()); env.elapsed = 0
    else:
        rnd_step()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
ms = e0.elapsed_time(e1) / steps
env.sim.enable_debug(True)                 # contact lists are only exported in debug mode: one more (untimed) step
if regime == "policy":
    env.step(env.act())
else:
    rnd_step()
torch.cuda.synchronize()
_, nc = env.sim.get_contacts()
print(f"{env_id} {n} envs regime {regime}: {ms:.3f} ms/step = {n / ms * 1e3:.4g} env-steps/s; contacts/env mean {float(nc.mean()):.2f}, "
      f"envs with contacts {float((nc > 0).mean()) * 100:.1f} %")
