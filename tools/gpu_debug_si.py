import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
np.set_printoptions(linewidth=220, precision=6, suppress=True)
from assistive_vr_gym_b200 import make, capi
from assistive_vr_gym_b200.envs import load_env_data
from oracle.oracle import Oracle, env_to_f64
from helpers import patch_blob
n = 256
env = make("ScratchItchJaco-v0", num_envs=n, device=0, seed=11)
env.reset()
st0 = env.get_state(); variant = env.variants.copy()
env.close()
blobs, resets = load_env_data("ScratchItchJaco.npz")
a = np.random.RandomState(0).uniform(-1, 1, (n, 7)).astype(np.float32)
for nsub in (1, 2, 3, 5):
    pb = [patch_blob(b, header={"substeps": nsub, "residual_thr": 0.0}) for b in blobs]
    oracles = [Oracle(b) for b in pb]
    sim = capi.Sim(n, 0)
    for v, b in enumerate(pb): sim.upload_model(v, b)
    sim.enable_debug(True)
    sim.set_state(st0, variant)
    obs = torch.zeros((n, 30), device="cuda"); rew = torch.zeros(n, device="cuda"); info = torch.zeros((n, 2), device="cuda")
    act = torch.as_tensor(a, device="cuda")
    sim.step(act.data_ptr(), obs.data_ptr(), rew.data_ptr(), 0, info.data_ptr(), 0)
    torch.cuda.synchronize()
    st = sim.get_state(); cont, nc = sim.get_contacts()
    print("substeps", nsub)
    for e in range(n):
        rec = env_to_f64(st0[e]).copy()
        o = oracles[int(variant[e])]
        oobs, orew, oinfo, oc = o.step(rec, a[e])
        dq = np.abs(rec[:32] - st[e, :32]).max()
        if dq > 1e-4:
            sh = o.model["shapes"]
            print(" env", e, "variant", variant[e], "dq %.3e" % dq, "overflow", st[e].view(np.int32)[166])
            print("  oracle", [(int(c[0]), int(c[1]), int(sh[int(c[0])]["type"]), int(sh[int(c[1])]["type"]), round(c[11], 5), round(c[12], 3), np.round(c[2:5], 4)) for c in oc])
            print("  gpu   ", [(int(c["shape_a"]), int(c["shape_b"]), round(float(c["dist"]), 5), round(float(c["force"]), 3), np.round(c["pos_a"], 4)) for c in cont[e, :nc[e]]])
    sim.close()
