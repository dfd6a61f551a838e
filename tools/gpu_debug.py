"""GPU-vs-oracle divergence report (writes gpurun_out/debug_parity.txt).  Usage: python tools/gpu_debug.py [n_env] [steps] [thr]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import make                 # noqa: E402
from oracle.oracle import Oracle, env_to_f64           # noqa: E402

n_env = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
env = make("ScratchItchJaco-v0", num_envs=n_env, device=0, seed=3)
env.sim.enable_debug(True)
obs0 = env.reset().cpu().numpy().copy()
state = env.get_state()
oracles = [Oracle(b) for b in env.blobs]
recs = [env_to_f64(state[e]).copy() for e in range(n_env)]
rng = np.random.RandomState(0)
lines = []
shown = 0
d0 = max(float(np.abs(oracles[int(env.variants[e])].reset_obs(recs[e]) - obs0[e]).max()) for e in range(n_env))
lines.append(f"reset obs max diff {d0:.3e}")
for t in range(steps):
    a = rng.uniform(-1, 1, (n_env, 7)).astype(np.float32)
    obs, rew, done, info = env.step(torch.as_tensor(a, device="cuda:0"))
    obs = obs.cpu().numpy(); rew = rew.cpu().numpy()
    st = env.get_state()
    cont, ncont = env.sim.get_contacts()
    dq = dqd = dobs = drew = 0.0
    mism = 0
    nc_tot = 0
    for e in range(n_env):
        o = oracles[int(env.variants[e])]
        oobs, orew, oinfo, ocont = o.step(recs[e], a[e])
        dq = max(dq, float(np.abs(recs[e][:32] - st[e, :32]).max()))
        dqd = max(dqd, float(np.abs(recs[e][32:64] - st[e, 32:64]).max()))
        dobs = max(dobs, float(np.abs(oobs - obs[e]).max()))
        drew = max(drew, abs(orew - float(rew[e])))
        gp = sorted((int(c["shape_a"]), int(c["shape_b"])) for c in cont[e, :ncont[e]])
        op = sorted((int(c[0]), int(c[1])) for c in ocont)
        nc_tot += len(op)
        derr = float(np.abs(recs[e][:64] - st[e, :64]).max())
        if t == 0 and len(op) > 0 and derr > 1e-3 and shown < 4:
            shown += 1
            lines.append(f"  == env {e} step 0 state err {derr:.3e} variant {int(env.variants[e])} iters gpu {int(st.view(np.int32)[e,167])}")
            for c in cont[e, :ncont[e]]:
                lines.append(f"     gpu  ({int(c['shape_a'])},{int(c['shape_b'])}) pa {np.round(c['pos_a'],5)} pb {np.round(c['pos_b'],5)} n {np.round(c['normal'],5)} d {c['dist']:.6f} f {c['force']:.5f}")
            for c in ocont:
                lines.append(f"     orcl ({int(c[0])},{int(c[1])}) pa {np.round(c[2:5],5)} pb {np.round(c[5:8],5)} n {np.round(c[8:11],5)} d {c[11]:.6f} f {c[12]:.5f}")
            lines.append(f"     dq  {np.round(recs[e][:24] - st[e, :24],4)}")
            lines.append(f"     dqd {np.round(recs[e][32:55] - st[e, 32:55],4)}")
        if gp != op:
            mism += 1
            lines.append(f"  step {t} env {e}: contact sets differ gpu={gp} oracle={op}")
    lines.append(f"step {t}: max|dq| {dq:.3e} max|dqd| {dqd:.3e} max|dobs| {dobs:.3e} max|drew| {drew:.3e} "
                 f"oracle contacts {nc_tot} set mismatches {mism} overflow {int(st.view(np.int32)[:, 166].max())}")
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/debug_parity.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
