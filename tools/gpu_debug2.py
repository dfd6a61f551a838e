import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from assistive_vr_gym_b200 import capi
from assistive_vr_gym_b200.envs import load_env_data
from assistive_vr_gym_b200.compiler.reset import sample_states
from oracle.oracle import Oracle, env_to_f64
from helpers import patch_blob
blobs, resets = load_env_data("ScratchItchJaco.npz")
pb = [patch_blob(b, header={"substeps": 1, "residual_thr": 0.0}) for b in blobs]
n = 256
env0, variant = sample_states(resets, n, np.random.RandomState(3))
sim = capi.Sim(n, 0)
for v, b in enumerate(pb): sim.upload_model(v, b)
sim.enable_debug(True); sim.set_state(env0, variant)
obs = torch.zeros((n, 30), device="cuda"); rew = torch.zeros(n, device="cuda"); info = torch.zeros((n, 2), device="cuda")
done = torch.zeros(n, dtype=torch.uint8, device="cuda")
a = np.random.RandomState(0).uniform(-1, 1, (n, 7)).astype(np.float32)
act = torch.as_tensor(a, device="cuda")
sim.step(act.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(), info.data_ptr(), 0)
torch.cuda.synchronize()
st = sim.get_state(); cont, nc = sim.get_contacts()
oracles = [Oracle(b) for b in pb]
np.set_printoptions(precision=5, suppress=True, linewidth=220)
for e in range(n):
    rec = env_to_f64(env0[e]).copy()
    o = oracles[int(variant[e])]
    oobs, orew, oinfo, ocont = o.step(rec, a[e])
    bad = any(abs(float(cg["force"]) - co[12]) > 5e-3 * max(1.0, abs(co[12])) for cg, co in zip(cont[e, :nc[e]], ocont))
    if bad:
        print("env", e, "variant", variant[e], "limit_scale", env0[e, 97], "q human", env0[e, 10:17])
        for c in cont[e, :nc[e]]:
            print("   gpu ", int(c['shape_a']), int(c['shape_b']), c['pos_a'], c['pos_b'], c['normal'], c['dist'], c['force'])
        for c in ocont:
            print("   orcl", int(c[0]), int(c[1]), c[2:5], c[5:8], c[8:11], c[11], c[12])
        print("   dq ", (rec[:24] - st[e, :24]))
        print("   dqd", (rec[32:55] - st[e, 32:55]))
        print("   qd orcl", rec[32:55])
