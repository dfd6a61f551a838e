"""e2e timing experiments: chunk counts of avg_step_host, raw copy bandwidth."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import make, capi
E = 196608
env = make("ScratchItchJaco-v0", num_envs=E, device=0, seed=1001)
env.reset()
a = capi.PinnedArray((E, 7), np.float32); a.array[...] = np.random.RandomState(0).uniform(-1, 1, (E, 7))
act_dev = torch.as_tensor(a.array, device="cuda")
for _ in range(3): env.step(act_dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): env.step(act_dev)
torch.cuda.synchronize()
print("device-only ms/step", (time.perf_counter() - t0) / 5 * 1e3)
for ch in (1, 2, 4, 8, 16):
    os.environ["AVG_CHUNKS"] = str(ch)
    env.step_host(a.array)
    t0 = time.perf_counter()
    for _ in range(5): env.step_host(a.array)
    print("chunks", ch, "e2e ms/step", (time.perf_counter() - t0) / 5 * 1e3)
# raw copies
obs_pin = torch.empty((E, 30), dtype=torch.float32).pin_memory()
for _ in range(2): obs_pin.copy_(env.obs); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): obs_pin.copy_(env.obs, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print("D2H 23.6 MB pinned: %.3f ms (%.1f GB/s)" % (dt * 1e3, E * 120 / dt / 1e9))
