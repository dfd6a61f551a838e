"""Parity statistics of the CUDA path against the oracle over many environments (profiles/parity_r2.json).

For every registered id: N environments are walked for 0..12 random env-steps on the GPU (so that states with contacts,
active joint limits and saturated motors are in the sample), then ONE env-step (frame_skip sub-steps) is taken from the
same float32 state by the CUDA path and by the float64 oracle, and compared: contact-pair sets of the last sub-step
(identical / differing only by pairs within 2e-6 of the contact threshold / differing), joint positions and velocities,
reward, total_force_on_human, task events (scratch / wipe counters).
usage: python tools/parity_report.py [n_env=1024] [ids...]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from assistive_vr_gym_b200 import make
from assistive_vr_gym_b200.envs import REGISTRY
from oracle.oracle import Oracle, env_to_f64, part_to_f64, part_masks

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ids = sys.argv[2:] or sorted(REGISTRY)
report = {}
for env_id in ids:
    env = make(env_id, num_envs=n, device=0, seed=21)
    env.sim.enable_debug(True)
    env.reset()
    variants = np.asarray(env.variants).copy()
    if env.has_particles and n > 48:          # Feeding / Drinking: the oracle needs ~1 s per Drinking env-step; a smaller sample
        env.close(); env = make(env_id, num_envs=48, device=0, seed=21); env.sim.enable_debug(True); env.reset(); variants = np.asarray(env.variants).copy()
    n_id = env.num_envs
    oracles = [Oracle(b) for b in env.blobs]
    na = env.sim.n_actions
    rng = np.random.RandomState(3)
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    # walk groups of environments for different numbers of steps: group k takes k random steps (others get zero actions and are reset to their state)
    n_save = n; n = n_id
    walk = rng.randint(0, 13, size=n)
    st = env.get_state()
    if not env.has_particles:
        for k in range(12):
            a = (torch.rand((n, na), device="cuda", generator=g) * 2 - 1)
            before = env.get_state()
            env.step(a); env.elapsed = 0
            after = env.get_state()
            keep = walk <= k                                      # environments that have finished their walk keep their state
            after[keep] = before[keep]
            env.set_state(after, variants)
    else:                                                         # particle ids: everybody walks 3 steps (no per-environment state mixing)
        for k in range(3):
            env.step(torch.rand((n, na), device="cuda", generator=g) * 2 - 1); env.elapsed = 0
    start = env.get_state()
    pstart = env.get_particles() if env.has_particles else None
    act = rng.uniform(-1, 1, (n, na)).astype(np.float32)
    obs, rew, done, info = env.step(torch.as_tensor(act, device="cuda"))
    torch.cuda.synchronize()
    st = env.get_state(); cont, nc = env.sim.get_contacts(); terms = env.sim.get_reward_terms()
    pend = env.get_particles() if env.has_particles else None
    rew = rew.cpu().numpy()
    events_equal = 0
    same = near = diff = with_contact = 0
    dq, dqd, dr, df, ev_equal = [], [], [], [], 0
    nq = int(oracles[0].model["header"]["n_jdof"])
    for e in range(n):
        o = oracles[int(variants[e])]
        rec = env_to_f64(start[e]).copy()
        prt = part_to_f64(pstart[e]).copy() if pstart is not None else None
        oobs, orew, oinfo, oc = o.step(rec, act[e], prt)
        if prt is not None:
            gp = part_to_f64(pend[e])
            events_equal += int(all(part_masks(gp, s_) == part_masks(prt, s_) for s_ in (576, 578, 584, 586, 588)))
        gp = [(int(c["shape_a"]), int(c["shape_b"])) for c in cont[e, :nc[e]]]
        op = [(int(c[0]), int(c[1])) for c in oc]
        if gp or op:
            with_contact += 1
        if gp == op:
            same += 1
        else:
            thr = lambda c: min(float(o.model["shapes"][int(c[0])]["thr"]), float(o.model["shapes"][int(c[1])]["thr"]))
            sym = set(gp) ^ set(op)
            border_o = all(abs(c[11] - thr(c)) < 2e-5 for c in oc if (int(c[0]), int(c[1])) in sym)
            border_g = all(abs(float(c["dist"]) - min(float(o.model["shapes"][int(c["shape_a"])]["thr"]), float(o.model["shapes"][int(c["shape_b"])]["thr"]))) < 2e-5
                           for c in cont[e, :nc[e]] if (int(c["shape_a"]), int(c["shape_b"])) in sym)
            if border_o and border_g:
                near += 1
            else:
                diff += 1
        dq.append(np.abs(rec[:nq] - st[e, :nq]).max()); dqd.append(np.abs(rec[32:32 + nq] - st[e, 32:32 + nq]).max())
        dr.append(abs(orew - rew[e])); df.append(abs(oinfo[0] - terms[e, 0]))
        ev_equal += int(rec[153] == st[e, 153] and [int(rec[170 + w]) for w in range(5)] == [int(x) for x in st[e].view(np.uint32)[170:175]])
    q = lambda v, p: float(np.percentile(np.asarray(v), p))
    report[env_id] = {"n_env": n_id, "with_contact": with_contact, "contact_sets_identical": same, "differ_only_at_threshold": near, "differ": diff,
                      "dq_median": q(dq, 50), "dq_p99": q(dq, 99), "dq_max": q(dq, 100), "dqd_median": q(dqd, 50), "dqd_p99": q(dqd, 99), "dqd_max": q(dqd, 100),
                      "dreward_p99": q(dr, 99), "dreward_max": q(dr, 100), "dforce_on_human_max": q(df, 100),
                      "task_counters_and_target_bitmaps_equal": ev_equal}
    if pstart is not None:
        report[env_id]["particle_alive_hit_event_masks_equal"] = events_equal
    n = n_save
    print(env_id, report[env_id], flush=True)
    env.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"note": "one env-step (5 sub-steps) from identical float32 states, CUDA float32 vs oracle float64; states sampled after 0-12 random env-steps",
           "ids": report}, open(os.path.join(ROOT, "gpurun_out", "parity_r2.json"), "w"), indent=1)
