import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
np.set_printoptions(linewidth=220, precision=6, suppress=True)
from assistive_vr_gym_b200 import capi
from assistive_vr_gym_b200.envs import load_env_data
from assistive_vr_gym_b200.compiler.reset import sample_states
from oracle.oracle import Oracle, env_to_f64
from helpers import patch_blob
blobs, resets = load_env_data("BedBathingJaco.npz")
n = 64
env0, variant = sample_states(resets, n, np.random.RandomState(5))
a = np.random.RandomState(0).uniform(-1, 1, (n, 7)).astype(np.float32)
hdr = {"substeps": 1, "residual_thr": 0.0}
cases = {"base": dict(header=hdr), "no friction": dict(header=hdr, friction=0.0), "no limits": dict(header=hdr, dof_flags_clear=1),
         "weak weld": dict(header=dict(hdr, weld_max_force=0.0)), "zero action": dict(header=hdr), "iters 1": dict(header=dict(hdr, solver_iters=1)),
         "iters 2": dict(header=dict(hdr, solver_iters=2)), "iters 5": dict(header=dict(hdr, solver_iters=5))}
for name, kw in cases.items():
    pb = [patch_blob(b, **kw) for b in blobs]
    oracles = [Oracle(b) for b in pb]
    sim = capi.Sim(n, 0)
    for v, b in enumerate(pb): sim.upload_model(v, b)
    sim.enable_debug(True)
    sim.set_state(env0, variant)
    obs = torch.zeros((n, 24), device="cuda"); rew = torch.zeros(n, device="cuda"); info = torch.zeros((n, 2), device="cuda")
    aa = a * 0 if name == "zero action" else a
    act = torch.as_tensor(aa, device="cuda")
    sim.step(act.data_ptr(), obs.data_ptr(), rew.data_ptr(), 0, info.data_ptr(), 0)
    torch.cuda.synchronize()
    st = sim.get_state(); cont, nc = sim.get_contacts()
    e = 29
    rec = env_to_f64(env0[e]).copy()
    oobs, orew, oinfo, oc = oracles[int(variant[e])].step(rec, aa[e])
    print(name, "force oracle", [round(c[12], 3) for c in oc], "gpu", [round(float(c["force"]), 3) for c in cont[e, :nc[e]]], "dqd %.3e" % np.abs(rec[32:64] - st[e, 32:64]).max())
    print("    qd oracle", rec[32:48]); print("    qd gpu   ", st[e, 32:48])
    sim.close()
