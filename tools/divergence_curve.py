"""Divergence-vs-horizon curve of the CUDA path against the CPU oracle (north_star: "with the divergence-vs-horizon
curve reported because contact dynamics are chaotic").  Writes gpurun_out/divergence_r2[_<id>].json.
usage: python tools/divergence_curve.py [n_env=256] [steps=100] [env_id=ScratchItchJaco-v0]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from assistive_vr_gym_b200 import make
from oracle.oracle import Oracle, env_to_f64

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
env_id = sys.argv[3] if len(sys.argv) > 3 else "ScratchItchJaco-v0"
env = make(env_id, num_envs=n, device=0, seed=11)
nq = int(np.frombuffer(env.blobs[0], dtype=np.int32, count=8)[7])           # n_jdof: the joint coordinates lead the record
env.sim.enable_debug(True)
env.reset()
st0 = env.get_state()
oracles = [Oracle(b) for b in env.blobs]
recs = [env_to_f64(st0[e]).copy() for e in range(n)]
rng = np.random.RandomState(0)
touched = np.zeros(n, dtype=bool)
rows = []
for t in range(T):
    a = rng.uniform(-1, 1, (n, 7)).astype(np.float32)
    obs, rew, done, info = env.step(torch.as_tensor(a, device="cuda")); env.elapsed = 0
    st = env.get_state(); cont, nc = env.sim.get_contacts(); rew = rew.cpu().numpy()
    dq = np.zeros(n); dr = np.zeros(n); same = np.zeros(n, dtype=bool)
    for e in range(n):
        oobs, orew, oinfo, oc = oracles[int(env.variants[e])].step(recs[e], a[e])
        if len(oc) or nc[e]:
            touched[e] = True
        dq[e] = np.abs(recs[e][:nq] - st[e, :nq]).max()
        dr[e] = abs(orew - rew[e])
        same[e] = sorted((int(c[0]), int(c[1])) for c in oc) == sorted((int(c["shape_a"]), int(c["shape_b"])) for c in cont[e, :nc[e]])
    def stats(mask):
        if not mask.any():
            return None
        return {"n": int(mask.sum()), "dq_median": float(np.median(dq[mask])), "dq_p90": float(np.percentile(dq[mask], 90)), "dq_max": float(dq[mask].max()),
                "dreward_median": float(np.median(dr[mask])), "dreward_max": float(dr[mask].max()), "contact_sets_equal": float(same[mask].mean())}
    rows.append({"env_step": t + 1, "contact_free_so_far": stats(~touched), "had_contact": stats(touched)})
    if (t + 1) in (1, 2, 5, 10, 20, 50, 100, 200):
        print(rows[-1])
out = {"env_id": env_id, "n_env": n, "steps": T, "note": "max |dq| over the %d joint coordinates, CUDA float32 vs oracle float64, random actions" % nq, "curve": rows}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
name = "divergence_r2.json" if env_id == "ScratchItchJaco-v0" else "divergence_r2_%s.json" % env_id[:-3]
json.dump(out, open(os.path.join(ROOT, "gpurun_out", name), "w"), indent=1)
