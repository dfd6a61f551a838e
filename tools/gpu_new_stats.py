"""Per-variant acceptance statistics of the `New` reset (avg_reset_new_kernel): usage gpu_new_stats.py <env_id> [n_env]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import make
env_id = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
env = make(env_id, num_envs=n, device=0, seed=5)
env.reset()
st = env.get_state(); v = np.asarray(env.variants)
gaps = st[:, 124 + 14]; att = st[:, 124 + 15]
for k in range(len(env.blobs)):
    s = v == k
    rd = env.reset_data[k]
    print(f"variant {k:2d}: h2m {float(rd['new_h2m']):.3f} waist {np.round(np.rad2deg(rd['new_waist']), 1)}  n {s.sum():4d}  clearance kept {100 * (gaps[s] >= 0.01 - 1e-6).mean():5.1f} %  "
          f"draws mean {att[s].mean():5.2f}  best clearance median {np.median(gaps[s]):+.4f} max {gaps[s].max():+.4f}")
