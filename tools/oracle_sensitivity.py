"""How sensitive is ONE env-step of the oracle to a 1e-7 perturbation of the joint positions (the size of float32 rounding)?
CPU-only measurement behind the with-contact tolerances of tests/test_gpu_parity.py: a step whose outcome changes by 1e-3 rad
under a 1e-7 rad perturbation (a joint-limit row that exists or not, a closest feature that switches, a friction direction taken
from a near-zero lateral velocity) cannot agree between a float32 and a float64 implementation, whatever the implementation.
Measured (ScratchItchJaco, 1500 environments after 5-40 drifting random steps, 311 in contact): contact-free max 2.4e-7 rad;
in contact median 9.5e-8, p90 2.8e-7, p99 1.7e-3, max 2.3e-3 rad -- the restated algorithm amplifies 1e-7 by up to 2e4 within
one env-step for ~1 % of the environments in contact.
usage: python tools/oracle_sensitivity.py [env npz] [n_env]"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from assistive_vr_gym_b200.envs import load_env_data
from assistive_vr_gym_b200.compiler.reset import sample_states
from oracle.oracle import Oracle, env_to_f64

name = sys.argv[1] if len(sys.argv) > 1 else "ScratchItchJaco.npz"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
modes = [0]
blobs, resets = load_env_data(name)
oracles = [Oracle(b) for b in blobs]
nq = int(oracles[0].model["header"]["n_jdof"]); na = oracles[0].n_act
for mode in modes:
    env0, variant = sample_states(resets, n, np.random.RandomState(3))
    rng = np.random.RandomState(5)
    walk = rng.randint(5, 40, size=n)
    dq = []; incontact = []; iters = []
    for e in range(n):
        o = oracles[int(variant[e])]
        rec = env_to_f64(env0[e]).copy()
        drift = rng.uniform(-1, 1, na) * 0.7
        for k in range(walk[e]):
            o.step(rec, np.clip(drift + rng.uniform(-1, 1, na) * 0.5, -1, 1))
        a = rng.uniform(-1, 1, na)
        r1 = rec.copy(); r2 = rec.copy()
        r2[:nq] += rng.uniform(-1, 1, nq) * 1e-7
        _, _, _, c1 = o.step(r1, a); _, _, _, c2 = o.step(r2, a)
        dq.append(np.abs(r1[:nq] - r2[:nq]).max()); incontact.append(len(c1) > 0 or len(c2) > 0); iters.append(r1[167])
    dq = np.array(dq); ic = np.array(incontact); iters = np.array(iters)
    f = lambda v: "n/a" if len(v) == 0 else f"median {np.median(v):.1e} p90 {np.percentile(v, 90):.1e} p99 {np.percentile(v, 99):.1e} max {v.max():.1e}"
    print(f"{ic.sum()} of {n} in contact | |dq| after one env-step from states 1e-7 apart: in contact {f(dq[ic])} | contact-free {f(dq[~ic])} | "
          f"solver iterations per env-step in contact {iters[ic].mean():.0f}, free {iters[~ic].mean():.0f}", flush=True)
