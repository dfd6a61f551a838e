"""Settle the human's right arm onto the mattress with the CUDA kernels (reference bed_bathing.py:284-292).

The reference drops the arm for 100 `p.stepSimulation()` calls at reset before it freezes the whole human.  The result
depends only on gender, so it is computed once, on a B200, by stepping the 'settle' worlds compiled by
tools/compile_models.py (data/BedBathingJacoSettle.npz) with zero robot action for 20 env-steps = 100 sub-steps, and
committed as data/bed_bathing_settle.json; tools/compile_models.py bakes that pose into the BedBathingJaco-v0 models.

    gpurun -- python tools/settle_bed_bathing.py --out gpurun_out/bed_bathing_settle.json
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import capi                                  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/bed_bathing_settle.json")
    ap.add_argument("--env-steps", type=int, default=20)                # x frame_skip 5 = 100 sub-steps, bed_bathing.py:291
    args = ap.parse_args()
    import torch
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(capi.__file__)), "data", "BedBathingJacoSettle.npz"))
    sim = capi.Sim(2, 0)
    for v in range(2):
        sim.upload_model(v, z[f"blob_{v}"].tobytes())
    sim.set_state(np.stack([z["init_0"], z["init_1"]]), np.array([0, 1], dtype=np.int32))
    dev = torch.device("cuda", 0)
    act = torch.zeros((2, sim.n_actions), dtype=torch.float32, device=dev)
    obs = torch.zeros((2, sim.n_obs), dtype=torch.float32, device=dev)
    rew = torch.zeros(2, dtype=torch.float32, device=dev)
    info = torch.zeros((2, 2), dtype=torch.float32, device=dev)
    trace = []
    for i in range(args.env_steps):
        sim.step(act.data_ptr(), obs.data_ptr(), rew.data_ptr(), 0, info.data_ptr(), 0)
        torch.cuda.synchronize()
        st = sim.get_state()
        trace.append([[float(st[v, q]) for q in z[f"arm_qidx_{v}"]] for v in range(2)])
    st = sim.get_state()
    out = {}
    for v, g in enumerate(("male", "female")):
        qi = z[f"arm_qidx_{v}"]
        out[g] = dict(arm_q=[float(st[v, q]) for q in qi], arm_qd=[float(st[v, 32 + q]) for q in qi],
                      overflow=int(st[v].view(np.int32)[166]), trace=[t[v] for t in trace])
        print(g, np.round(out[g]["arm_q"], 5), "max |qd|", np.abs(out[g]["arm_qd"]).max(), "overflow", out[g]["overflow"])
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", args.out)
