"""Generate tests/golden/scratch_itch_jaco_oracle.npz: initial env records, actions, and the CPU oracle's outputs.

There are no reference golden vectors for this path and PyBullet cannot be imported here, so this
fixture pins OUR oracle (regression anchor for both the oracle and the CUDA path), not the reference.  Re-run after an
intentional change of the restated algorithm:  python tools/make_golden.py
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200.envs import load_env_data
from assistive_vr_gym_b200.compiler.reset import sample_states
from oracle.oracle import Oracle, env_to_f64

for name, n_act in (("ScratchItchJaco", 7), ("ScratchItchJacoHuman", 17)):
    blobs, resets = load_env_data(name + ".npz")
    n, T = 12, 12
    env, var = sample_states(resets, n, np.random.RandomState(20261018))
    acts = np.random.RandomState(7).uniform(-1.2, 1.2, (T, n, n_act)).astype(np.float32)   # some entries beyond the clip range
    oracles = [Oracle(b) for b in blobs]
    n_obs = oracles[0].n_obs
    obs0 = np.zeros((n, n_obs)); obs = np.zeros((T, n, n_obs)); rew = np.zeros((T, n)); info = np.zeros((T, n, 8))
    states = np.zeros((T, n, 192)); ncont = np.zeros((T, n), dtype=np.int32)
    for e in range(n):
        rec = env_to_f64(env[e]).copy()
        o = oracles[int(var[e])]
        obs0[e] = o.reset_obs(rec)
        for t in range(T):
            obs[t, e], rew[t, e], info[t, e], c = o.step(rec, acts[t, e])
            states[t, e] = rec; ncont[t, e] = len(c)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f"{name}_oracle.npz")
    np.savez_compressed(out, env=env, variant=var, actions=acts, obs0=obs0, obs=obs, reward=rew, info=info, states=states, ncontacts=ncont)
    print("wrote", out, "contacts per step", ncont.sum(1))
