"""Compile the reference assets into ModelBlobs + reset pools and store them under assistive_vr_gym_b200/data/.

Run in the build container (needs /root/reference); the outputs are derived data (hull vertices, inertias, IK start
poses) that travel to the GPU box, where the reference tree does not exist.

    python tools/compile_models.py [--assets /root/reference/assistive_gym/envs/assets] [--pool 64]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200.compiler.blob import scene_to_blob          # noqa: E402
from assistive_vr_gym_b200.compiler import xform as X                  # noqa: E402
from assistive_vr_gym_b200.compiler.reset import (build_reset_data, build_reset_data_bed_bathing, bed_bathing_settle_record,  # noqa: E402
                                                  toc_search_jaco, toc_search, ik_dls)
from assistive_vr_gym_b200.compiler.scene import (build_scratch_itch, build_bed_bathing, load_settled_arm_q,  # noqa: E402
                                                  urdf_to_multibody, load_robot)


def compile_bed_bathing_new(assets: str, out_dir: str, robot_type: str, n_var: int, attempts: int):
    """BedBathing<Robot>New-v0 (__init__.py:122-134): `n_var` variants per gender, each with its own person (height, waist pose)
    and robot base pose (task-oriented-configuration search against that person's shoulder / elbow / wrist at the preset arm
    pose, bed_bathing.py:303-318).  No settle stage: the New branch starts from the preset arm pose (bed_bathing.py:269)."""
    rng = np.random.RandomState(4001)
    payload = {}
    name = "BedBathing" + {"jaco": "Jaco", "pr2": "PR2"}[robot_type] + "New"
    if robot_type == "jaco":
        robot = urdf_to_multibody(os.path.join(assets, "jaco", "j2s7s300_gym.urdf"), 0, "jaco")
    else:
        robot, rs = load_robot(assets, "pr2")
    v = 0
    for gender in ("male", "female"):
        for k in range(n_var):
            h2m, waist, pkw, acc = draw_feasible_person(rng, gender, lambda kw: build_bed_bathing(assets, robot_type, gender, stage="play", **kw),
                                                        lambda sc: build_reset_data_bed_bathing(sc, np.zeros(7)))
            probe = build_bed_bathing(assets, robot_type, gender, stage="play", **pkw)
            cf = probe.multibodies[1].com_frames(probe.q_human_reset)
            goals = [cf[9][0], cf[11][0], cf[13][0]]
            if robot_type == "jaco":
                xy, yaw, q_start, reached = toc_search_jaco(robot, [1, 2, 3, 4, 5, 6, 7], np.array([-0.5, -0.1, 1.0]),
                                                            X.quat_from_euler([0, np.pi / 2.0, 0]), goals, rng, [0.1, 0.55, 0.6], attempts=attempts)
            else:
                xy, yaw, q_start, reached = toc_search(robot, rs["arm"], np.array([-0.5, -0.1, 1.0]), X.quat_from_euler([0, 0, 0]), goals, rng,
                                                       [0.0, 0.0, 0.0], attempts=attempts, random_position=0.5, ee_link=rs["ee_link"])
            scene = build_bed_bathing(assets, robot_type, gender, stage="play", base_xy_yaw=(float(xy[0]), float(xy[1]), float(yaw)), **pkw)
            blob = scene_to_blob(scene)
            payload[f"blob_{v}"] = np.frombuffer(blob, dtype=np.uint8)
            rd = build_reset_data_bed_bathing(scene, q_start)
            rd.update(new_mode=np.asarray(1), hum_jitter=np.asarray(np.deg2rad(10.0)), new_min_dist=np.asarray(0.01),
                      new_h2m=np.asarray(h2m), new_waist=np.asarray(waist), frozen_mask=np.asarray(scene.frozen_mask))
            for key, a in rd.items():
                payload[f"reset_{v}_{key}"] = a
            print(name, gender, k, "h2m", round(h2m, 3), "waist", np.round(np.rad2deg(waist), 1), "arm draws accepted", round(acc, 2),
                  "base", np.round(xy, 3), "yaw", round(float(yaw), 3), "goals reached", reached,
                  scene.info["n_pairs"], "pairs", scene.info["n_target"], "targets", len(blob), "bytes")
            v += 1
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **payload)
    print("wrote", name + ".npz")


def compile_bed_bathing(assets: str, out_dir: str, n_base: int, attempts: int):
    """BedBathingJaco-v0 in two stages.  Stage 1 (always): the 'settle' worlds of both genders + their start records ->
    BedBathingJacoSettle.npz; tools/settle_bed_bathing.py runs them for 100 sub-steps on a GPU and writes the settled
    arm pose to data/bed_bathing_settle.json.  Stage 2 (when that file exists): per gender `n_base` robot base poses from
    the task-oriented-configuration search (env.py:486-585), one play variant each -> BedBathingJaco.npz."""
    payload = {}
    for v, gender in enumerate(("male", "female")):
        scene = build_bed_bathing(assets, "jaco", gender, stage="settle")
        blob = scene_to_blob(scene)
        payload[f"blob_{v}"] = np.frombuffer(blob, dtype=np.uint8)
        payload[f"init_{v}"] = bed_bathing_settle_record(scene)
        payload[f"arm_qidx_{v}"] = np.asarray([b.qidx for b in scene.bodies if b.art == 1], dtype=np.int32)
        print("settle", gender, scene.info["n_pairs"], "pairs", len(blob), "bytes")
    np.savez_compressed(os.path.join(out_dir, "BedBathingJacoSettle.npz"), **payload)
    print("wrote BedBathingJacoSettle.npz")
    settled = load_settled_arm_q()
    if settled["male"] is None or settled["female"] is None:
        print("no data/bed_bathing_settle.json yet: run tools/settle_bed_bathing.py on a GPU box, then re-run this tool")
        return
    payloads = {False: {}, True: {}}
    rng = np.random.RandomState(1001)
    robot = urdf_to_multibody(os.path.join(assets, "jaco", "j2s7s300_gym.urdf"), 0, "jaco")
    v = 0
    for gender in ("male", "female"):
        probe = build_bed_bathing(assets, "jaco", gender, stage="play", arm_q=settled[gender])
        cf = probe.multibodies[1].com_frames(probe.q_human_reset)
        goals = [cf[9][0], cf[11][0], cf[13][0]]                              # shoulder, elbow, wrist: bed_bathing.py:303-305,325
        for k in range(n_base):
            xy, yaw, q_start, reached = toc_search_jaco(robot, [1, 2, 3, 4, 5, 6, 7], np.array([-0.5, -0.1, 1.0]),
                                                        X.quat_from_euler([0, np.pi / 2.0, 0]), goals, rng, [0.1, 0.55, 0.6],
                                                        attempts=attempts)
            for human_control in (False, True):                               # BedBathingJaco-v0 / BedBathingJacoHuman-v0 share the base poses
                scene = build_bed_bathing(assets, "jaco", gender, human_control=human_control, stage="play", arm_q=settled[gender],
                                          base_xy_yaw=(float(xy[0]), float(xy[1]), float(yaw)))
                blob = scene_to_blob(scene)
                payloads[human_control][f"blob_{v}"] = np.frombuffer(blob, dtype=np.uint8)
                rd = build_reset_data_bed_bathing(scene, q_start)
                for key, a in rd.items():
                    payloads[human_control][f"reset_{v}_{key}"] = a
                print("play", gender, k, "human_control" if human_control else "", "base", np.round(xy, 3), "yaw", round(float(yaw), 3),
                      "goals reached", reached, scene.info["n_pairs"], "pairs", len(blob), "bytes")
            v += 1
    np.savez_compressed(os.path.join(out_dir, "BedBathingJaco.npz"), **payloads[False])
    np.savez_compressed(os.path.join(out_dir, "BedBathingJacoHuman.npz"), **payloads[True])
    print("wrote BedBathingJaco.npz, BedBathingJacoHuman.npz")


def draw_new_human(rng: np.random.RandomState, gender: str):
    """Per-episode draws of the `New` ids that this build bakes per model variant: hipbone_to_mouth_height
    (scratch_itch.py:158) and the three waist angles (scratch_itch.py:211)."""
    h2m = rng.uniform(0.6 - 0.1, 0.6 + 0.1) if gender == "male" else rng.uniform(0.54 - 0.1, 0.54 + 0.1)
    waist = rng.uniform(np.deg2rad(-10), np.deg2rad(10), size=3)
    return float(h2m), tuple(float(w) for w in waist)


def draw_feasible_person(rng, gender, build_probe, reset_data_of, min_accept: float = 0.25):
    """Draw (height, waist pose) until the per-episode arm resampling has a fair chance for that person: the reference redraws the
    waist together with the arm until the pose is collision-free (scratch_itch.py:198-223), so persons whose arm cannot be placed
    never appear in its episodes.  build_probe(pkw) -> scene, reset_data_of(scene) -> reset data with the robot at a start pose."""
    from assistive_vr_gym_b200.compiler.reset import new_variant_acceptance
    for attempt in range(40):
        h2m, waist = draw_new_human(rng, gender)
        pkw = dict(new=True, hipbone_to_mouth_height=h2m, waist=waist)
        scene = build_probe(pkw)
        rd = reset_data_of(scene)
        rd.update(hum_jitter=np.deg2rad(10.0), new_min_dist=0.01)
        acc = new_variant_acceptance(scene_to_blob(scene), rd, np.random.RandomState(attempt))
        if acc >= min_accept:
            return h2m, waist, pkw, acc
    raise RuntimeError("no feasible person drawn")


def compile_scratch_itch_jaco_new(assets: str, out_dir: str, n_var: int, pool: int):
    """ScratchItchJacoNew-v0 (__init__.py:45-50): `n_var` model variants per gender, each with its own drawn height and waist
    pose (the reference draws both per episode); the arm pose is drawn per episode on the device (AvgResetTable.new_mode)."""
    payload = {}
    rng = np.random.RandomState(2001)
    v = 0
    for gender in ("male", "female"):
        for k in range(n_var):
            h2m, waist, pkw, acc = draw_feasible_person(rng, gender, lambda kw: build_scratch_itch(assets, "jaco", gender, **kw),
                                                        lambda sc: build_reset_data(sc, np.random.RandomState(7), ik_pool=1))
            scene = build_scratch_itch(assets, "jaco", gender, **pkw)
            blob = scene_to_blob(scene)
            payload[f"blob_{v}"] = np.frombuffer(blob, dtype=np.uint8)
            rd = build_reset_data(scene, rng, ik_pool=pool)
            rd.update(new_mode=np.asarray(1), hum_jitter=np.asarray(np.deg2rad(10.0)), new_min_dist=np.asarray(0.01),
                      new_h2m=np.asarray(h2m), new_waist=np.asarray(waist))
            for key, a in rd.items():
                payload[f"reset_{v}_{key}"] = a
            print("ScratchItchJacoNew", gender, k, "h2m", round(h2m, 3), "waist", np.round(np.rad2deg(waist), 1), "arm draws accepted", round(acc, 2),
                  scene.info["n_pairs"], "pairs", len(blob), "bytes")
            v += 1
    np.savez_compressed(os.path.join(out_dir, "ScratchItchJacoNew.npz"), **payload)
    print("wrote ScratchItchJacoNew.npz")


def compile_pr2(assets: str, out_dir: str, task: str, n_base: int, attempts: int, pool: int, new: bool = False):
    """<Task>PR2-v0 / <Task>PR2Human-v0 for ScratchItch and BedBathing: per gender `n_base` base poses from the
    task-oriented-configuration search (env.py:486-585 as called at scratch_itch.py:245 / bed_bathing.py:318: left arm,
    tool link 76, random_position 0.5 m, +-30 deg), one model variant each (the PR2's static branches are baked into the
    world at that pose).  ScratchItch draws its start target per episode (scratch_itch.py:243): each variant carries a
    pool of IK start poses for `pool` target draws reached from its base."""
    settled = load_settled_arm_q()
    rng = np.random.RandomState(1001)
    robot, rs = load_robot(assets, "pr2")
    joints = rs["arm"]
    lower = np.array([robot.links[j].lower for j in joints]); upper = np.array([robot.links[j].upper for j in joints])
    ik_lo = np.where(lower > upper, -2 * np.pi, lower); ik_hi = np.where(lower > upper, 2 * np.pi, upper)
    start_quat = X.quat_from_euler([0, 0, 0])
    payloads = {False: {}, True: {}}
    name = {"scratch_itch": "ScratchItchPR2", "bed_bathing": "BedBathingPR2"}[task] + ("New" if new else "")
    assert not new or task == "scratch_itch"
    new_kw = {}

    def build(gender, human_control, base):
        if task == "scratch_itch":
            return build_scratch_itch(assets, "pr2", gender, human_control=human_control, base_xy_yaw=base, **new_kw)
        return build_bed_bathing(assets, "pr2", gender, human_control=human_control, stage="play", arm_q=settled[gender], base_xy_yaw=base)

    v = 0
    for gender in ("male", "female"):
        pos_offset = [0.1, 0.0, 0.0] if task == "scratch_itch" else [0.0, 0.0, 0.0]
        for k in range(n_base):
            if new:                                                            # ScratchItchPR2New-v0: the variant's own height and waist pose
                h2m, waist, new_kw, acc = draw_feasible_person(
                    rng, gender, lambda kw: build_scratch_itch(assets, "pr2", gender, **kw),
                    lambda sc: build_reset_data_bed_bathing(sc, np.zeros(7)))
                print(name, gender, k, "h2m", round(h2m, 3), "waist", np.round(np.rad2deg(waist), 1), "arm draws accepted", round(acc, 2))
            probe = build(gender, False, (0.0, 0.0, 0.0))
            cf = probe.multibodies[1].com_frames(probe.q_human_reset)
            goals = [cf[9][0], cf[11][0], cf[13][0]]                              # shoulder, elbow, wrist
            if task == "scratch_itch":
                start_pos = np.array([-0.55, 0, 0.8]) + rng.uniform(-0.05, 0.05, size=3)   # scratch_itch.py:243
            else:
                start_pos = np.array([-0.5, -0.1, 1.0])                                   # bed_bathing.py:315
            xy, yaw, q_start, reached = toc_search(robot, joints, start_pos, start_quat, goals, rng, pos_offset, attempts=attempts,
                                                   random_position=0.5, ee_link=rs["ee_link"])
            starts = [q_start]
            tries = 0
            while task == "scratch_itch" and len(starts) < pool and tries < 20 * pool:    # further episodes from the same base
                tries += 1
                tp = np.array([-0.55, 0, 0.8]) + rng.uniform(-0.05, 0.05, size=3)
                q, ep, eq = ik_dls(robot, rs["ee_link"], joints, ik_lo, ik_hi, tp, start_quat, rng.uniform(ik_lo, ik_hi), iters=200)
                if ep < 0.03 and eq < 0.03:
                    starts.append(q)
            for human_control in ((False,) if new else (False, True)):
                scene = build(gender, human_control, (float(xy[0]), float(xy[1]), float(yaw)))
                blob = scene_to_blob(scene)
                payloads[human_control][f"blob_{v}"] = np.frombuffer(blob, dtype=np.uint8)
                rd = build_reset_data_bed_bathing(scene, np.asarray(starts))
                if new:
                    rd.update(new_mode=np.asarray(1), hum_jitter=np.asarray(np.deg2rad(10.0)), new_min_dist=np.asarray(0.01),
                              new_h2m=np.asarray(h2m), new_waist=np.asarray(waist))
                for key, a in rd.items():
                    payloads[human_control][f"reset_{v}_{key}"] = a
                print(name, gender, k, "human_control" if human_control else "", "base", np.round(xy, 3), "yaw", round(float(yaw), 3),
                      "goals reached", reached, "start poses", len(starts), scene.info["n_pairs"], "pairs", scene.info["n_mshape"],
                      "moving shapes", len(blob), "bytes")
            v += 1
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **payloads[False])
    if not new:
        np.savez_compressed(os.path.join(out_dir, name + "Human.npz"), **payloads[True])
    print("wrote", name + ".npz" + ("" if new else ", " + name + "Human.npz"))


def compile_feeding_drinking(assets: str, out_dir: str, task: str, robot: str, n_base: int, attempts: int, pool: int, new: bool = False):
    """Feeding<Robot>[Human]-v0 / Drinking<Robot>[Human]-v0 (feeding.py:144-331, drinking.py:159-335).  Jaco: fixed base, one
    variant per gender.  PR2 (feeding.py:266-270) and the build-defined Sawyer / Baxter ids: per gender `n_base` base poses from
    the task-oriented-configuration search with the start target and the mouth as goals, one variant each.
    `new` = <Task><Robot>New-v0 (__init__.py:206-218,290-302; Jaco and PR2): `n_base` variants per gender, each with its own
    person (height and waist pose, drawn per episode in the reference: feeding.py:171,233) and, on the PR2, its own base pose."""
    from assistive_vr_gym_b200.compiler.scene_fd import build_feeding_drinking, ROBOT_FD, start_target
    from assistive_vr_gym_b200.compiler.reset_fd import build_reset_data_fd
    rng = np.random.RandomState(1001 if not new else 3001)
    rec = ROBOT_FD[(task, robot)]
    name = ({"feeding": "Feeding", "drinking": "Drinking"}[task] + {"jaco": "Jaco", "pr2": "PR2", "sawyer": "Sawyer", "baxter": "Baxter"}[robot]
            + ("New" if new else ""))
    payloads = {False: {}, True: {}}
    v = 0
    for gender in ("male", "female"):
        n_var = n_base if (new or rec["toc"] is not None) else 1
        for k in range(n_var):
            pkw = {}
            if new:
                h2m, waist = draw_new_human(rng, gender)
                pkw = dict(new=True, hipbone_to_mouth_height=h2m, waist=waist)
            base = (0.0, 0.0, 0.0)
            if rec["toc"] is not None:
                probe = build_feeding_drinking(assets, task, robot, gender, **pkw)
                hp, hq = probe.multibodies[1].com_frames(probe.q_human_reset)[27]
                mouth = hp + X.quat_rotate(hq, np.asarray(probe.header["task_f"][19:22], float))     # feeding.py:253-256
                robot_mb, rs = load_robot(assets, robot, arm="right")
                centre, quat = start_target(task, robot)
                sp = centre + rng.uniform(-0.05, 0.05, size=3)
                xy, yaw, q_start, reached = toc_search(robot_mb, rs["arm"], sp, quat, [mouth, mouth], rng, rec["toc"], attempts=attempts,
                                                       random_position=0.5, ee_link=rs["ee_link"])
                base = (float(xy[0]), float(xy[1]), float(yaw))
                print(name, gender, "base", k, np.round(xy, 3), "yaw", round(float(yaw), 3), "goals reached", reached)
            for human_control in ((False,) if new else (False, True)):
                scene = build_feeding_drinking(assets, task, robot, gender, human_control=human_control, base_xy_yaw=base, **pkw)
                blob = scene_to_blob(scene)
                payloads[human_control][f"blob_{v}"] = np.frombuffer(blob, dtype=np.uint8)
                rd = build_reset_data_fd(scene, np.random.RandomState(1001 + v), ik_pool=pool)
                if new:
                    rd.update(new_mode=np.asarray(1), hum_jitter=np.asarray(0.0), new_min_dist=np.asarray(0.01),
                              new_h2m=np.asarray(h2m), new_waist=np.asarray(waist))
                for key, a in rd.items():
                    payloads[human_control][f"reset_{v}_{key}"] = a
                print(name, gender, "human_control" if human_control else "", scene.info["n_body"], "bodies", scene.info["n_dof"], "dof",
                      scene.info["n_mshape"], "moving shapes", scene.info["n_cshape"], "compound children", scene.info["n_pairs"], "pairs", len(blob), "bytes")
            v += 1
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **payloads[False])
    if not new:
        np.savez_compressed(os.path.join(out_dir, name + "Human.npz"), **payloads[True])
    print("wrote", name + ".npz" + ("" if new else ", " + name + "Human.npz"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--assets", default="/root/reference/assistive_gym/envs/assets")
    ap.add_argument("--pool", type=int, default=64)
    ap.add_argument("--only", default="", help="scratch_itch | bed_bathing | pr2 | scratch_itch_pr2 | bed_bathing_pr2 (default: all)")
    ap.add_argument("--bases", type=int, default=8, help="BedBathing: robot base poses (model variants) per gender")
    ap.add_argument("--attempts", type=int, default=100, help="BedBathing: base poses tried per TOC search (env.py:486)")
    ap.add_argument("--fd-bases", type=int, default=4, help="Feeding / Drinking on PR2 / Sawyer / Baxter: robot base poses (model variants) per gender")
    args = ap.parse_args()
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assistive_vr_gym_b200", "data")
    os.makedirs(out_dir, exist_ok=True)
    if args.only in ("", "bed_bathing"):
        compile_bed_bathing(args.assets, out_dir, args.bases, args.attempts)
    if args.only in ("", "pr2", "scratch_itch_pr2"):
        compile_pr2(args.assets, out_dir, "scratch_itch", args.bases, args.attempts, min(args.pool, 16))
    if args.only in ("", "new", "scratch_itch_jaco_new"):
        compile_scratch_itch_jaco_new(args.assets, out_dir, args.bases, min(args.pool, 8))
    if args.only in ("", "new", "scratch_itch_pr2_new"):
        compile_pr2(args.assets, out_dir, "scratch_itch", args.bases, args.attempts, 4, new=True)
    for robot in ("jaco", "pr2"):
        if args.only in ("", "new", f"bed_bathing_{robot}_new"):
            compile_bed_bathing_new(args.assets, out_dir, robot, args.bases, args.attempts)
    if args.only in ("", "pr2", "bed_bathing_pr2"):
        compile_pr2(args.assets, out_dir, "bed_bathing", args.bases, args.attempts, 1)
    for task in ("feeding", "drinking"):
        for robot in ("jaco", "pr2", "sawyer", "baxter"):
            if args.only in ("", "fd", task, f"{task}_{robot}"):
                compile_feeding_drinking(args.assets, out_dir, task, robot, args.fd_bases, args.attempts, min(args.pool, 8))
            if robot in ("jaco", "pr2") and args.only in ("", "new", "fd_new", f"{task}_{robot}_new"):
                compile_feeding_drinking(args.assets, out_dir, task, robot, args.bases, args.attempts, min(args.pool, 8), new=True)
    for human_control in ((False, True) if args.only in ("", "scratch_itch") else ()):
        payload = {}
        rng = np.random.RandomState(1001)              # env.py:53 default seed
        for v, gender in enumerate(("male", "female")):
            scene = build_scratch_itch(args.assets, "jaco", gender, human_control=human_control)
            blob = scene_to_blob(scene)
            payload[f"blob_{v}"] = np.frombuffer(blob, dtype=np.uint8)
            rd = build_reset_data(scene, rng, ik_pool=args.pool)
            for k, a in rd.items():
                payload[f"reset_{v}_{k}"] = a
            print(gender, "human_control" if human_control else "", scene.info["n_pairs"], "pairs", len(blob), "bytes",
                  {k: round(e, 5) for k, e in scene.info["hull_errors"].items()})
        name = "ScratchItchJaco" + ("Human" if human_control else "") + ".npz"
        np.savez_compressed(os.path.join(out_dir, name), **payload)
        print("wrote", name)
