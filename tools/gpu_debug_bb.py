import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
np.set_printoptions(linewidth=220, precision=6, suppress=True)
from assistive_vr_gym_b200 import make
from assistive_vr_gym_b200.envs import load_env_data
from oracle.oracle import Oracle, env_to_f64
env_id = "BedBathingJacoHuman-v0"
oracles = [Oracle(b) for b in load_env_data(env_id[:-3] + ".npz")[0]]
n = 64
env = make(env_id, num_envs=n, device=0, seed=5); env.sim.enable_debug(True)
env.reset(); st0 = env.get_state()
recs = [env_to_f64(st0[e]).copy() for e in range(n)]
rng = np.random.RandomState(0)
sh = oracles[0].model["shapes"]
for t in range(10):
    a = rng.uniform(-1, 1, (n, 17)).astype(np.float32)
    obs, rew, done, info = env.step(torch.as_tensor(a, device="cuda"))
    st = env.get_state(); cont, nc = env.sim.get_contacts(); rew = rew.cpu().numpy()
    dq = np.zeros(n); dqd = np.zeros(n); dr = np.zeros(n); same = np.zeros(n, bool); kinds = []
    for e in range(n):
        o = oracles[int(env.variants[e])]
        oobs, orew, oinfo, oc = o.step(recs[e], a[e])
        dq[e] = np.abs(recs[e][:32] - st[e, :32]).max(); dqd[e] = np.abs(recs[e][32:64] - st[e, 32:64]).max(); dr[e] = abs(orew - rew[e])
        op = [(int(c[0]), int(c[1])) for c in oc]; gp = [(int(c["shape_a"]), int(c["shape_b"])) for c in cont[e, :nc[e]]]
        same[e] = op == gp
        if e == 0 and t < 2:
            print("  env0 contacts", [(x, int(o.model["shapes"][x[0]]["ref_body"]), int(o.model["shapes"][x[1]]["ref_body"]), round(c[11], 5), round(c[12], 4)) for x, c in zip(op, oc)])
    print("step", t, "dq med %.2e p90 %.2e max %.2e | dqd med %.2e max %.2e | drew max %.2e | sets equal %.2f | ncont mean %.1f" % (
        np.median(dq), np.percentile(dq, 90), dq.max(), np.median(dqd), dqd.max(), dr.max(), same.mean(), nc.mean()))
