"""Development aid: state of Feeding / Drinking environments right after the device reset (IK error, residual velocities)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200 import make
from assistive_vr_gym_b200.compiler.blob import read_blob
env_id = sys.argv[1] if len(sys.argv) > 1 else "FeedingJaco-v0"; n = int(sys.argv[2]) if len(sys.argv) > 2 else 6; seed = int(sys.argv[3]) if len(sys.argv) > 3 else 13
env = make(env_id, num_envs=n, device=0, seed=seed)
env.reset()
st = env.get_state(); pt = env.get_particles()
m = read_blob(env.blobs[0]); h = m["header"]
td = int(m["bodies"][int(h["tool_body"])]["dof"]); nj = int(h["n_jdof"]); npart = int(h["n_particle"])
for e in range(n):
    sp = np.abs(pt[e, 192:384]).reshape(3, 64)[:, :npart].max(0)
    print(e, "variant", env.variants[e], "ik err %.5f" % st[e, 148], "tool |v| %.4f |w| %.4f" % (np.linalg.norm(st[e, 32 + td:35 + td]), np.linalg.norm(st[e, 35 + td:38 + td])),
          "max |qd| joints %.4f" % np.abs(st[e, 32:32 + nj]).max(), "particle speeds", np.round(sp, 3), "z", np.round(pt[e, 128:128 + npart] - st[e, int(m["bodies"][int(h["tool_body"])]["qidx"]) + 2], 3),
          "tremor", st[e, 99], "frozen", st[e].view(np.uint32)[175], "ovf", st[e].view(np.int32)[166], pt[e].view(np.int32)[591])
    print("   q-mtarget", np.round(st[e, :nj] - st[e, 64:64 + nj], 4))
# contacts of the environments that did not come to rest
env.sim.enable_debug(True)
env.step(torch.zeros((n, env.sim.n_actions), device="cuda"))
cont, cnt = env.sim.get_contacts()
names = {0: "robot", 1: "human", 2: "tool", 3: "furniture", 4: "plane", 5: "table", 6: "bowl", 7: "particle"}
for e in range(n):
    if np.abs(st[e, 32:32 + nj]).max() > 0.01:
        sh = read_blob(env.blobs[int(env.variants[e])])["shapes"]
        print("env", e, "contacts:", [(names[int(sh[c["shape_a"]]["ref_body"])], int(sh[c["shape_a"]]["ref_link"]), names[int(sh[c["shape_b"]]["ref_body"])],
                                       int(sh[c["shape_b"]]["ref_link"]), round(float(c["dist"]), 4), round(float(c["force"]), 2)) for c in cont[e, :cnt[e]]])
        print("    arm q", np.round(st[e, :7], 3), "tool pos", np.round(st[e, int(m["bodies"][int(h["tool_body"])]["qidx"]):][:3], 3), "bowl", np.round(st[e, 124:127], 3))
