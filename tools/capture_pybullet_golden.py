"""Capture golden trajectories from the UNMODIFIED reference (assistive_gym + its PyBullet fork).

This is the tool SURVEY.md 8c asks for: the reference's arithmetic for the hot path lives in the third-party `pybullet`
module (reference setup.py:18), which is absent from this build container, so parity is UNPINNED until a file produced by
this script exists.  Run it on any machine where `import assistive_gym` works (python 3.6 era: gym, pybullet fork,
keras 2.3 / tensorflow 1.14 -- see the reference's README.md:26-27), then commit the output under tests/golden/:

    python tools/capture_pybullet_golden.py --env ScratchItchJaco-v0 --seed 1001 --steps 10 \
        --out tests/golden/pybullet_ScratchItchJaco-v0_seed1001.npz

tests/test_pybullet_golden.py picks every tests/golden/pybullet_*.npz up and compares the oracle (and, on a GPU, the CUDA
path) against it; without such a file the test is skipped with the reason "parity unpinned".

What is recorded (everything the comparison needs, nothing from our own code):
  model:    per body (robot, human, tool): getNumJoints, getJointInfo (type, limits, damping, axis, parent frame),
            getDynamicsInfo (mass, local inertia diagonal, inertial frame, friction), getCollisionShapeData;
            getPhysicsEngineParameters()
  episode:  gender, robot base pose, impairment parameters, target_on_arm / wiping targets, task-specific bodies
  per sub-step (hooked p.stepSimulation, i.e. frame_skip entries per env-step, env.py:341-349):
            joint positions / velocities of every movable joint of robot and human, base pose + twist of the tool,
            getContactPoints() of the tool and of the robot: (bodyA, bodyB, linkA, linkB, posA, posB, normal, distance, force)
  per env-step: action, observation, reward, info['total_force_on_human'], info['task_success']
  Feeding / Drinking (round 2): the bowl pose, and per sub-step the position and linear velocity of every food / water sphere
            (NaN once the reference has removed it from self.foods / self.waters); per env-step how many are left and how many
            have hit the person (feeding.py:92-121, drinking.py:95-136) -- what the particle-event parity needs
  BedBathing: per env-step the number of wiping targets left on the upper arm / forearm (bed_bathing.py:111-125)
  `New` ids:  hipbone_to_mouth_height and every human joint angle at reset (waist draw, arm draw: scratch_itch.py:158,211-217)
"""
import argparse
import json

import numpy as np


def body_model(p, body, cid):
    out = {"num_joints": p.getNumJoints(body, physicsClientId=cid), "joints": [], "dynamics": [], "shapes": []}
    for j in range(-1, out["num_joints"]):
        if j >= 0:
            ji = p.getJointInfo(body, j, physicsClientId=cid)
            out["joints"].append({"index": ji[0], "name": ji[1].decode(), "type": ji[2], "damping": ji[6], "friction": ji[7],
                                  "lower": ji[8], "upper": ji[9], "max_force": ji[10], "max_velocity": ji[11], "link": ji[12].decode(),
                                  "axis": list(ji[13]), "parent_pos": list(ji[14]), "parent_orn": list(ji[15]), "parent": ji[16]})
        di = p.getDynamicsInfo(body, j, physicsClientId=cid)
        out["dynamics"].append({"link": j, "mass": di[0], "lateral_friction": di[1], "inertia_diag": list(di[2]),
                                "inertial_pos": list(di[3]), "inertial_orn": list(di[4]), "restitution": di[5],
                                "rolling_friction": di[6], "spinning_friction": di[7], "contact_damping": di[8], "contact_stiffness": di[9]})
        for cs in p.getCollisionShapeData(body, j, physicsClientId=cid):
            out["shapes"].append({"link": cs[1], "geom": cs[2], "dims": list(cs[3]), "file": cs[4].decode() if isinstance(cs[4], bytes) else cs[4],
                                  "pos": list(cs[5]), "orn": list(cs[6])})
    return out


def movable(p, body, cid):
    return [j for j in range(p.getNumJoints(body, physicsClientId=cid)) if p.getJointInfo(body, j, physicsClientId=cid)[2] != p.JOINT_FIXED]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="ScratchItchJaco-v0")
    ap.add_argument("--seed", type=int, default=1001)          # the constructor default, env.py:53
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()

    import gym
    import assistive_gym  # noqa: F401  (registers the ids, assistive_gym/__init__.py)
    import pybullet as p

    env = gym.make(args.env)
    env.seed(args.seed)
    obs0 = env.reset()
    u = env.unwrapped
    cid = u.id
    tool = getattr(u, "tool", None) or getattr(u, "spoon", None) or getattr(u, "cup", None)
    bodies = {"robot": u.robot, "human": u.human, "tool": tool}
    mov = {k: movable(p, b, cid) for k, b in bodies.items()}

    particles0 = list(getattr(u, "foods", None) or getattr(u, "waters", None) or [])      # body ids in creation order (feeding.py:300-307)

    def particle_state():
        alive = set(getattr(u, "foods", None) or getattr(u, "waters", None) or [])
        out = np.full((len(particles0), 6), np.nan)
        for i, b in enumerate(particles0):
            if b in alive:
                pos, _ = p.getBasePositionAndOrientation(b, physicsClientId=cid)
                lin, _ = p.getBaseVelocity(b, physicsClientId=cid)
                out[i] = list(pos) + list(lin)
        return out.tolist()

    substeps = []
    real_step = p.stepSimulation

    def hooked(*a, **kw):
        r = real_step(*a, **kw)
        rec = {}
        for k, b in bodies.items():
            js = p.getJointStates(b, mov[k], physicsClientId=cid) if mov[k] else []
            rec[k + "_q"] = [s[0] for s in js]; rec[k + "_qd"] = [s[1] for s in js]
        pos, orn = p.getBasePositionAndOrientation(tool, physicsClientId=cid)
        lin, ang = p.getBaseVelocity(tool, physicsClientId=cid)
        rec["tool_base"] = list(pos) + list(orn) + list(lin) + list(ang)
        cps = []
        for body in (tool, u.robot):
            for c in p.getContactPoints(bodyA=body, physicsClientId=cid):
                cps.append([c[1], c[2], c[3], c[4]] + list(c[5]) + list(c[6]) + list(c[7]) + [c[8], c[9]])
        rec["contacts"] = cps
        if particles0:
            rec["particles"] = particle_state()
        substeps.append(rec)
        return r

    meta = {"env": args.env, "seed": args.seed, "gender": u.gender, "robot_type": u.robot_type,
            "physics": {k: (v if not isinstance(v, tuple) else list(v)) for k, v in p.getPhysicsEngineParameters(physicsClientId=cid).items()},
            "model": {k: body_model(p, b, cid) for k, b in bodies.items()}, "movable": mov,
            "robot_base": [list(x) for x in p.getBasePositionAndOrientation(u.robot, physicsClientId=cid)],
            "human_base": [list(x) for x in p.getBasePositionAndOrientation(u.human, physicsClientId=cid)],
            "human_impairment": getattr(u.world_creation, "human_impairment", None),
            "human_strength": float(getattr(u.world_creation, "human_strength", 1.0)),
            "human_limit_scale": float(getattr(u.world_creation, "human_limit_scale", 1.0)),
            "human_tremors": [float(x) for x in np.atleast_1d(getattr(u.world_creation, "human_tremors", []))],
            "target_on_arm": [float(x) for x in np.atleast_1d(getattr(u, "target_on_arm", []))],
            "limb": int(getattr(u, "limb", -1)),
            "hipbone_to_mouth_height": float(getattr(u, "hipbone_to_mouth_height", 0.0) or 0.0), "new": bool(getattr(u, "new", False)),
            "human_q_all": [s[0] for s in p.getJointStates(u.human, list(range(p.getNumJoints(u.human, physicsClientId=cid))), physicsClientId=cid)],
            "bowl": ([list(x) for x in p.getBasePositionAndOrientation(u.bowl, physicsClientId=cid)] if hasattr(u, "bowl") else None),
            "n_particles": len(particles0),
            "total_target_count": int(getattr(u, "total_target_count", 0) or 0),
            "target_human_joint_positions": [float(x) for x in np.atleast_1d(getattr(u, "target_human_joint_positions", []))]}
    init = {}
    for k, b in bodies.items():
        js = p.getJointStates(b, mov[k], physicsClientId=cid) if mov[k] else []
        init[k + "_q"] = [s[0] for s in js]; init[k + "_qd"] = [s[1] for s in js]
    pos, orn = p.getBasePositionAndOrientation(tool, physicsClientId=cid)
    init["tool_base"] = list(pos) + list(orn) + [0.0] * 6

    p.stepSimulation = hooked
    n_act = env.action_space.shape[0]
    actions = np.random.RandomState(0).uniform(-1, 1, (args.steps, n_act)).astype(np.float32)
    obs, rew, force, success, marks, left, hit, targets_left = [], [], [], [], [], [], [], []
    for t in range(args.steps):
        o, r, d, info = env.step(actions[t])
        obs.append(np.asarray(o, dtype=np.float64)); rew.append(float(r))
        force.append(float(info["total_force_on_human"])); success.append(int(info["task_success"]))
        marks.append(len(substeps))
        left.append(len(getattr(u, "foods", None) or getattr(u, "waters", None) or []))
        hit.append(len(getattr(u, "foods_hit_person", []) or []))
        targets_left.append([len(getattr(u, "targets_upperarm", []) or []), len(getattr(u, "targets_forearm", []) or [])])
    p.stepSimulation = real_step

    def stack(key):
        return np.asarray([s[key] for s in substeps], dtype=np.float64)

    max_c = max([len(s["contacts"]) for s in substeps] + [1])
    contacts = np.full((len(substeps), max_c, 15), np.nan)
    for i, s in enumerate(substeps):
        for j, c in enumerate(s["contacts"]):
            contacts[i, j] = c
    np.savez_compressed(args.out, meta=json.dumps(meta), init=json.dumps(init), actions=actions, obs0=np.asarray(obs0, dtype=np.float64),
                        obs=np.asarray(obs), reward=np.asarray(rew), total_force_on_human=np.asarray(force), task_success=np.asarray(success),
                        substep_marks=np.asarray(marks), robot_q=stack("robot_q"), robot_qd=stack("robot_qd"), human_q=stack("human_q"),
                        human_qd=stack("human_qd"), tool_base=stack("tool_base"), contacts=contacts,
                        particles=(stack("particles") if particles0 else np.zeros((0, 0, 6))), particles_left=np.asarray(left),
                        particles_hit_person=np.asarray(hit), targets_left=np.asarray(targets_left))
    print("wrote", args.out, "substeps", len(substeps), "max contacts per sub-step", max_c)


if __name__ == "__main__":
    main()
