"""Per-dof divergence between the CUDA path and the oracle on contact-free PR2 environments (development aid)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import make
from assistive_vr_gym_b200.envs import load_env_data, REGISTRY
from oracle.oracle import Oracle, env_to_f64
env_id = sys.argv[1] if len(sys.argv) > 1 else "BedBathingPR2-v0"
blobs = load_env_data(REGISTRY[env_id]["data"])[0]
n, T = 64, 10
env = make(env_id, num_envs=n, device=0, seed=5); env.sim.enable_debug(True)
if os.environ.get("NO_EARLY_EXIT"):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from helpers import patch_blob
    blobs = [patch_blob(b, header={"residual_thr": 0.0}) for b in blobs]
    for v, b in enumerate(blobs):
        env.sim.upload_model(v, b)
oracles = [Oracle(b) for b in blobs]
env.reset()
st0 = env.get_state()
recs = [env_to_f64(st0[e]).copy() for e in range(n)]
clean = np.ones(n, dtype=bool)
rng = np.random.RandomState(0)
na = env.sim.n_actions
for t in range(T):
    a = rng.uniform(-1, 1, (n, na)).astype(np.float32)
    env.step(torch.as_tensor(a, device="cuda"))
    st = env.get_state(); cont, ncont = env.sim.get_contacts()
    dq = np.zeros(32); dqd = np.zeros(32)
    for e in range(n):
        oobs, orew, oinfo, oc = oracles[int(env.variants[e])].step(recs[e], a[e])
        if len(oc) or ncont[e]:
            clean[e] = False
        if clean[e]:
            dq = np.maximum(dq, np.abs(recs[e][:32] - st[e, :32])); dqd = np.maximum(dqd, np.abs(recs[e][32:64] - st[e, 32:64]))
    print("t", t, "clean", clean.sum(), "dq", np.array2string(dq[:18], precision=1, floatmode="maxprec"), "\n   dqd", np.array2string(dqd[:18], precision=1))
