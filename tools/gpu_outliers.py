"""Where do the largest one-env-step deviations between the CUDA path and the oracle come from?  Walks N environments for
0..K policy / random steps, takes ONE env-step from identical float32 states on both sides and prints, for the worst
environments, the contact lists of both sides (pair, distance, force) and the joints that differ.
usage: python tools/gpu_outliers.py [env_id] [n_env] [top] [walk max]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from assistive_vr_gym_b200 import make
from oracle.oracle import Oracle, env_to_f64

env_id = sys.argv[1] if len(sys.argv) > 1 else "ScratchItchJaco-v0"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
top = int(sys.argv[3]) if len(sys.argv) > 3 else 6
wmax = int(sys.argv[4]) if len(sys.argv) > 4 else 25
env = make(env_id, num_envs=n, device=0, seed=21)
env.sim.enable_debug(True)
env.reset()
variants = np.asarray(env.variants).copy()
oracles = [Oracle(b) for b in env.blobs]
na = env.sim.n_actions
rng = np.random.RandomState(3)
g = torch.Generator(device="cuda"); g.manual_seed(4)
walk = rng.randint(0, wmax, size=n)
for k in range(wmax - 1):
    a = (torch.rand((n, na), device="cuda", generator=g) * 2 - 1)
    before = env.get_state()
    env.step(a); env.elapsed = 0
    after = env.get_state()
    keep = walk <= k
    after[keep] = before[keep]
    env.set_state(after, variants)
start = env.get_state()
act = rng.uniform(-1, 1, (n, na)).astype(np.float32)
env.step(torch.as_tensor(act, device="cuda")); torch.cuda.synchronize()
st = env.get_state(); cont, nc = env.sim.get_contacts()
nq = int(oracles[0].model["header"]["n_jdof"])
dq = np.zeros(n); recs = []; ocs = []
for e in range(n):
    rec = env_to_f64(start[e]).copy()
    _, _, _, oc = oracles[int(variants[e])].step(rec, act[e])
    recs.append(rec); ocs.append(oc)
    dq[e] = np.abs(rec[:nq] - st[e, :nq]).max()
hadc = np.array([len(ocs[e]) > 0 or nc[e] > 0 for e in range(n)])
print(f"{env_id}: {n} envs, {hadc.sum()} with contact at the end of the step; |dq| with contact: median {np.median(dq[hadc]):.2e} p90 {np.percentile(dq[hadc], 90):.2e} "
      f"p99 {np.percentile(dq[hadc], 99):.2e} max {dq[hadc].max():.2e}; without: max {dq[~hadc].max():.2e}")
print("fraction of in-contact envs above 1e-5 / 1e-4 / 1e-3 / 1e-2:", [(dq[hadc] > t).mean().round(4) for t in (1e-5, 1e-4, 1e-3, 1e-2)])
shapes = oracles[0].model["shapes"]
def nm(o, i):
    s = o.model["shapes"][int(i)]
    return f"{int(i)}(b{int(s['ref_body'])}/l{int(s['ref_link'])}/t{int(s['type'])})"
for e in np.argsort(-dq)[:top]:
    o = oracles[int(variants[e])]
    print(f"--- env {e}: |dq| {dq[e]:.3e}, walked {walk[e]} steps, variant {variants[e]}")
    d = recs[e][:nq] - st[e, :nq]
    print("   dq per joint:", np.array2string(d, precision=4, suppress_small=True))
    print("   q0 -> q (oracle) change:", np.array2string(recs[e][:nq] - start[e, :nq], precision=3, suppress_small=True))
    print("   oracle contacts:", [(nm(o, c[0]), nm(o, c[1]), round(float(c[11]), 5), round(float(c[12]), 3)) for c in ocs[e]])
    print("   cuda   contacts:", [(nm(o, c['shape_a']), nm(o, c['shape_b']), round(float(c['dist']), 5), round(float(c['force']), 3)) for c in cont[e, :nc[e]]])
    print("   solver iterations (cuda, 5 sub-steps):", int(st[e].view(np.int32)[167]), " overflow", int(st[e].view(np.int32)[166]))
