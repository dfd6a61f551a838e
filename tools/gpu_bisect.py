"""Bisect GPU-vs-oracle differences by patching model header fields (substeps, iterations, friction)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import capi
from assistive_vr_gym_b200.envs import load_env_data
from assistive_vr_gym_b200.compiler.blob import HEADER_DT, SHAPE_DT
from assistive_vr_gym_b200.compiler.reset import sample_states
from oracle.oracle import Oracle, env_to_f64

blobs, resets = load_env_data("ScratchItchJaco.npz")
n = 64
env0, variant = sample_states(resets, n, np.random.RandomState(3))

def patched(blob, substeps, iters, friction, thr):
    b = bytearray(blob)
    h = np.frombuffer(b, dtype=HEADER_DT, count=1)
    h["substeps"] = substeps; h["solver_iters"] = iters; h["residual_thr"] = thr
    if friction is not None:
        sh = np.frombuffer(b, dtype=SHAPE_DT, count=int(h["n_shape"][0]), offset=int(h["off_shape"][0]))
        sh["friction"] = friction
    return bytes(b)

for (ss, it, fr, thr) in [(1, 50, None, 1e-7), (1, 1, None, 0.0), (1, 2, None, 0.0), (1, 50, 0.0, 0.0), (1, 50, None, 0.0), (5, 50, 0.0, 0.0)]:
    pb = [patched(b, ss, it, fr, thr) for b in blobs]
    sim = capi.Sim(n, 0)
    for v, b in enumerate(pb): sim.upload_model(v, b)
    sim.enable_debug(True)
    sim.set_state(env0, variant)
    obs = torch.zeros((n, 30), device="cuda"); rew = torch.zeros(n, device="cuda"); info = torch.zeros((n, 2), device="cuda")
    done = torch.zeros(n, dtype=torch.uint8, device="cuda")
    a = np.random.RandomState(0).uniform(-1, 1, (n, 7)).astype(np.float32)
    act = torch.as_tensor(a, device="cuda")
    sim.step(act.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(), info.data_ptr(), 0)
    torch.cuda.synchronize()
    st = sim.get_state(); cont, nc = sim.get_contacts()
    oracles = [Oracle(b) for b in pb]
    worst = []
    for e in range(n):
        rec = env_to_f64(env0[e]).copy()
        oobs, orew, oinfo, ocont = oracles[int(variant[e])].step(rec, a[e])
        err = float(np.abs(rec[:64] - st[e, :64]).max())
        worst.append((err, e, len(ocont), int(nc[e])))
    worst.sort(reverse=True)
    wc = max([w for w in worst if w[2] > 0] or [(0, 0, 0, 0)])
    wn = max([w for w in worst if w[2] == 0 and w[3] == 0] or [(0, 0, 0, 0)])
    print(f"substeps {ss} iters {it} friction {fr} thr {thr}: worst with contacts {wc}, worst without {wn}")
    if ss == 1 and it == 1:
        e = wc[1]
        rec = env_to_f64(env0[e]).copy()
        oobs, orew, oinfo, ocont = oracles[int(variant[e])].step(rec, a[e])
        np.set_printoptions(precision=6, suppress=True, linewidth=200)
        for c in cont[e, :nc[e]]:
            print("   gpu ", int(c['shape_a']), int(c['shape_b']), c['pos_a'], c['pos_b'], c['normal'], c['dist'], c['force'])
        for c in ocont:
            print("   orcl", int(c[0]), int(c[1]), c[2:5], c[5:8], c[8:11], c[11], c[12])
        print("   dqd", rec[32:55] - st[e, 32:55])
    sim.close()
