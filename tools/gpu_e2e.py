"""e2e timing experiments: chunk counts of avg_step_host (each config from a fresh reset, same actions)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import make, capi
E = 196608
env = make("ScratchItchJaco-v0", num_envs=E, device=0, seed=1001)
acts = [capi.PinnedArray((E, 7), np.float32) for _ in range(13)]
for i, a in enumerate(acts): a.array[...] = np.random.RandomState(i).uniform(-1, 1, (E, 7))
dev = [torch.as_tensor(a.array, device="cuda") for a in acts]
env.reset()
for k in range(3): env.step(dev[k])
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(3, 13): env.step(dev[k])
torch.cuda.synchronize()
print("device-only ms/step", (time.perf_counter() - t0) / 10 * 1e3)
for ch in (1, 2, 4, 8):
    os.environ["AVG_CHUNKS"] = str(ch)
    env.seed(1001); env.reset()
    for k in range(3): env.step_host(acts[k].array)
    t0 = time.perf_counter()
    for k in range(3, 13): env.step_host(acts[k].array)
    print("chunks", ch, "e2e ms/step", (time.perf_counter() - t0) / 10 * 1e3)
