"""Device reset cost with the pool start poses and with the on-device IK (development aid)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import make
n = int(sys.argv[1]) if len(sys.argv) > 1 else 393216
for env_id in ("ScratchItchJaco-v0", "ScratchItchPR2-v0"):
    for ik in (False, True):
        env = make(env_id, num_envs=n, device=0, seed=1, device_ik=ik)
        env.reset_device(seed=5); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(5):
            env.reset_device(seed=5)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        st = env.get_state()
        err = st[:, 127] if ik else np.zeros(1)
        print(f"{env_id} device_ik={ik}: reset of {n} envs {dt * 1e3:.2f} ms = {n / dt:.3e} resets/s"
              + (f", IK position error median {np.median(err):.2e} p99 {np.percentile(err, 99):.2e}, within 0.03: {(err < 0.03).mean() * 100:.2f} %" if ik else ""), flush=True)
        env.close()
