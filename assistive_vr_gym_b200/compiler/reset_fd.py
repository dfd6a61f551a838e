"""Episode initial states for Feeding / Drinking (CPU, numpy): the random part of reference `FeedingEnv.reset`
(`feeding.py:171-185,242-245,276-280,291-310`) and `DrinkingEnv.reset` (`drinking.py:185-201,241-243,272-276,291-311`).

Per episode the reference draws: gender (:172), impairment and its parameters (world_creation.py:66-72), tremor amplitudes
for the 4 head joints (+-20 deg, world_creation.py:138-139), the head angles 25..27 (+-30 deg, feeding.py:243), the bowl
position (+-0.05 in x and y, feeding.py:184), the start target of the end effector (+-0.05 cube around a point above the
bowl, feeding.py:276; around (-0.2, -0.5, 1) for Drinking, drinking.py:272) -> IK with random restarts (util.py:34-57), the
tool in the gripper (world_creation.py:331-337), the particle grid above the tool (feeding.py:301-305), then 100 calls of
stepSimulation to drop the particles into the tool (feeding.py:318-320) -- that settle is run by the simulator itself
(C-ABI `avg_settle`), or by the oracle in CPU-only tests; this module only produces the pre-settle state.

The host sampler draws from a pool of complete, consistent draws (bowl offset + start target + arm pose); the device sampler
(csrc `avg_reset_fd_kernel`) draws afresh and solves the IK on the GPU.
"""
from __future__ import annotations

from typing import List

import numpy as np

from . import xform as X
from .blob import ENV_STRIDE
from .reset import (E_EBODY, E_HUMAN_KP, E_LIMIT_SCALE, E_MTARGET, E_Q, E_STRENGTH, E_TARGET_H, E_TREMOR, E_TREMOR_ON, ik_dls)
from .scene import CompiledScene
from .scene_fd import BOWL_CENTER, BOWL_QUAT, ROBOT_FD, particle_grid, start_target

E_FROZEN = 175
P_STRIDE, P_POS, P_VEL, P_ANG, P_ALIVE = 592, 0, 192, 384, 576


def tool_pose_for(scene: CompiledScene, q_arm) -> np.ndarray:
    """World pose (pos, quat) of the tool's body frame when the arm is at `q_arm` and the tool sits in the gripper
    (`init_tool`, world_creation.py:331-337)."""
    robot = scene.multibodies[0]
    spec = scene.robot_spec
    qrob = dict(spec["q_preset"])
    qrob.update({j: float(q_arm[k]) for k, j in enumerate(scene.robot_arm_joints)})
    for j, sg in zip(spec["fingers"], scene.finger_signs):
        qrob[j] = float(scene.finger_open) * sg
    ee_p, ee_q = robot.com_frames(qrob)[spec["ee_link"]]
    base_p, base_q = X.tf_mul(ee_p, ee_q, *scene.tool_offset)
    at = scene.attach[2][-1]
    ip, iq = X.tf_inv(at.pos, at.quat)
    bp, bq = X.tf_mul(base_p, base_q, ip, iq)
    return np.concatenate([bp, bq])


def pose_hits_table(scene: CompiledScene, q_arm, margin: float = 0.01) -> bool:
    """Does a robot link's box overlap the table top at this arm pose?  The reference's `ik_random_restarts(step_sim=True)`
    (util.py:41-52) steps the simulation 5 times after every IK attempt and keeps the attempt only if the end effector is still
    on its target, which weeds out poses that start inside the table (it pushes the arm away); the host pool applies the same
    idea geometrically (the device reset runs the reference's test itself, csrc avg_reset_check_kernel)."""
    if scene.task != "feeding":
        return False
    from .scene_fd import TABLE_POS
    robot = scene.multibodies[0]
    spec = scene.robot_spec
    qrob = dict(spec["q_preset"])
    qrob.update({j: float(q_arm[k]) for k, j in enumerate(scene.robot_arm_joints)})
    lf = robot.link_frames(qrob)
    lo_t = TABLE_POS + np.array([-0.75, -0.5, 0.675]) - margin; hi_t = TABLE_POS + np.array([0.75, 0.5, 0.725]) + margin   # table_tall.urdf: box 1.5 x 1 x 0.05 at z = 0.7
    arm_links = set(scene.robot_arm_joints) | set(spec["fingers"]) | {spec["ee_link"]}
    for l in robot.links:
        if l.ref_index not in arm_links or not l.shapes:
            continue
        p, r = lf[l.ref_index]
        for s in l.shapes:
            sp, sq = X.tf_mul(p, r, s.pos, s.quat)
            R = X.quat_to_mat(sq)
            slo, shi = s.local_aabb()
            c = R @ (0.5 * (slo + shi)) + sp; hh = np.abs(R) @ (0.5 * (shi - slo))
            if np.all(c + hh >= lo_t) and np.all(c - hh <= hi_t):
                return True
    return False


def build_reset_data_fd(scene: CompiledScene, rng: np.random.RandomState, ik_pool: int = 32) -> dict:
    robot, human = scene.multibodies[0], scene.multibodies[1]
    task, rtype = scene.task, scene.robot_type
    rec = ROBOT_FD[(task, rtype)]
    joints = scene.robot_arm_joints
    spec = scene.robot_spec
    lower = np.array([robot.links[j].lower for j in joints]); upper = np.array([robot.links[j].upper for j in joints])
    ik_lo = np.where(lower > upper, -2 * np.pi, lower); ik_hi = np.where(lower > upper, 2 * np.pi, upper)     # util.py:86-88
    centre, target_quat = start_target(task, rtype)
    tol = float(rec["tol"])
    pool_q, pool_tool, pool_bowl, pool_target = [], [], [], []
    tries = 0
    while len(pool_q) < ik_pool:
        tries += 1
        bowl_off = np.array([rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), 0.0]) if task == "feeding" else np.zeros(3)   # feeding.py:184
        target_pos = centre + bowl_off + rng.uniform(-0.05, 0.05, size=3)                                                     # feeding.py:276
        best, best_err, ok = None, np.inf, False
        for _ in range(40):                                            # max_ik_random_restarts
            rest = rng.uniform(np.maximum(ik_lo, -np.pi), np.minimum(ik_hi, np.pi))
            q, ep, eq = ik_dls(robot, spec["ee_link"], joints, ik_lo, ik_hi, target_pos, target_quat, rest, iters=200)
            if ep < tol and eq < tol and not pose_hits_table(scene, q):    # random_restart_threshold (+ the 5-step test, see pose_hits_table)
                best, ok = q, True
                break
            if ep < best_err:
                best, best_err = q, ep
        if not ok and tries < 8 * ik_pool:
            continue                                                   # keep the pool to poses that reached their target
        pool_q.append(np.asarray(best)); pool_tool.append(tool_pose_for(scene, best))
        pool_bowl.append(bowl_off); pool_target.append(target_pos)
    arm_qidx, arm_dof, fin_qidx, fin_dof, fin_q, hum_qidx, hum_dof, hum_joint = [], [], [], [], [], [], [], []
    tool_qidx = -1
    for b in scene.bodies:
        if b.art == 0 and b.jtype != 2:
            if b.ref_joint in joints:
                arm_qidx.append(b.qidx); arm_dof.append(b.dof)
            else:
                fin_qidx.append(b.qidx); fin_dof.append(b.dof)
                fin_q.append(float(scene.finger_open) * scene.finger_signs[spec["fingers"].index(b.ref_joint)])
        elif b.art == 1:
            hum_qidx.append(b.qidx); hum_dof.append(b.dof); hum_joint.append(b.ref_joint)
        elif b.art == 2:
            tool_qidx = b.qidx
    ii = lambda a: np.asarray(a, dtype=np.int64)
    # the end-effector link's COM frame in the frame of its dynamic body, for the device IK (weld parent = EE frame o tool offset)
    wb, wp, wq = scene.frames[2]
    ip, iq = X.tf_inv(*scene.tool_offset)
    ee_p, ee_q = X.tf_mul(np.asarray(wp, float), np.asarray(wq, float), ip, iq)
    return dict(pool_q=np.asarray(pool_q), pool_tool=np.asarray(pool_tool), pool_bowl=np.asarray(pool_bowl), pool_target=np.asarray(pool_target),
                arm_qidx=ii(arm_qidx), arm_dof=ii(arm_dof), fin_qidx=ii(fin_qidx), fin_dof=ii(fin_dof), fin_q=np.asarray(fin_q, dtype=np.float64),
                hum_qidx=ii(hum_qidx), hum_dof=ii(hum_dof), hum_joint=ii(hum_joint),
                hum_lower=np.array([human.links[j].lower for j in hum_joint], dtype=np.float64),
                hum_upper=np.array([human.links[j].upper for j in hum_joint], dtype=np.float64),
                hum_reset=np.array([scene.q_human_reset.get(j, 0.0) for j in hum_joint], dtype=np.float64),
                tool_qidx=np.asarray(tool_qidx), limb_dims=np.zeros((2, 2)), human_control=np.asarray(int(scene.human_control)),
                task=np.asarray(int(scene.header["task"])), n_target=np.asarray(0), fin_open=np.asarray(float(scene.finger_open)),
                n_particle=np.asarray(int(scene.header["n_particle"])), grid=particle_grid(task),
                head_mask=np.asarray(int(scene.header["head_frozen_mask"])),
                bowl_center=BOWL_CENTER.copy(), bowl_quat=BOWL_QUAT.copy(), has_bowl=np.asarray(int(task == "feeding")),
                ik_ee_body=np.asarray(int(wb)), ik_ee_pos=ee_p, ik_ee_quat=ee_q, ik_center=centre, ik_quat=target_quat,
                ik_tol=np.asarray(tol))


def sample_states_fd(reset_data: List[dict], n: int, rng: np.random.RandomState, genders: np.ndarray | None = None):
    """Pre-settle env records and particle records of n environments (see module docstring).
    -> (env [n, ENV_STRIDE] f32, part [n, P_STRIDE] f32, variant [n] i32)"""
    env = np.zeros((n, ENV_STRIDE), dtype=np.float32)
    env_u = env.view(np.uint32)
    part = np.zeros((n, P_STRIDE), dtype=np.float32)
    part_u = part.view(np.uint32)
    nv = len(reset_data)
    npg = max(nv // 2, 1)
    gender = (np.asarray(genders, dtype=np.int32) if genders is not None else rng.randint(min(nv, 2), size=n).astype(np.int32))
    impairment = rng.randint(4, size=n)                        # 0 none, 1 limits, 2 weakness, 3 tremor (world_creation.py:67)
    limit_scale = np.where(impairment == 1, rng.uniform(0.5, 1.0, size=n), 1.0)
    strength = np.where(impairment == 2, rng.uniform(0.25, 1.0, size=n), 1.0)
    tremor = rng.uniform(np.deg2rad(-20), np.deg2rad(20), size=(n, 4)) * (impairment == 3)[:, None]      # world_creation.py:138-139
    head = rng.uniform(np.deg2rad(-30), np.deg2rad(30), size=(n, 3))                                   # feeding.py:243 (joints 25, 26, 27)
    pool_pick = rng.randint(1 << 30, size=n)
    variant = (gender * npg + pool_pick % npg).astype(np.int32)
    for v in range(nv):
        idx = np.nonzero(variant == v)[0]
        if idx.size == 0:
            continue
        rd = reset_data[v]
        k = (pool_pick[idx] // npg) % len(rd["pool_q"])
        qa = rd["pool_q"][k]
        ls = limit_scale[idx][:, None]
        hq = np.tile(rd["hum_reset"][None, :], (idx.size, 1))
        for c, j in enumerate((25, 26, 27)):
            hq[:, list(rd["hum_joint"]).index(j)] = head[idx, c]
        hq = np.clip(hq, rd["hum_lower"][None, :] * ls, rd["hum_upper"][None, :] * ls)      # enforce_joint_limits, world_creation.py:172
        env[np.ix_(idx, E_Q + rd["arm_qidx"])] = qa
        env[np.ix_(idx, E_MTARGET + rd["arm_dof"])] = qa
        env[np.ix_(idx, E_Q + rd["fin_qidx"])] = rd["fin_q"][None, :]
        env[np.ix_(idx, E_MTARGET + rd["fin_dof"])] = rd["fin_q"][None, :]
        env[np.ix_(idx, E_Q + rd["hum_qidx"])] = hq
        env[np.ix_(idx, E_MTARGET + rd["hum_dof"])] = hq
        env[np.ix_(idx, E_TARGET_H + rd["hum_joint"] - 24)] = hq                            # target_human_joint_positions, feeding.py:248
        tq = int(rd["tool_qidx"])
        tool = rd["pool_tool"][k]
        env[idx, E_Q + tq:E_Q + tq + 7] = tool
        if int(rd["has_bowl"]):
            env[idx, E_EBODY:E_EBODY + 3] = rd["bowl_center"][None, :] + rd["pool_bowl"][k]
            env[idx, E_EBODY + 3:E_EBODY + 7] = rd["bowl_quat"][None, :]
        active = bool(rd["human_control"]) | (impairment[idx] == 3)                         # feeding.py:244
        env_u[idx, E_FROZEN] = np.where(active, 0, int(rd["head_mask"])).astype(np.uint32)
        npart = int(rd["n_particle"])
        # particles on the reference's grid relative to the tool's base position: the tool's COM frame origin
        # (getBasePositionAndOrientation, feeding.py:292); the tool's inertial offset is zero, so that is the body frame origin
        pos = tool[:, None, :3] + rd["grid"][None, :, :]
        for c in range(3):
            part[np.ix_(idx, P_POS + 64 * c + np.arange(npart))] = pos[:, :, c]
        mask = (1 << npart) - 1
        part_u[idx, P_ALIVE] = np.uint32(mask & 0xffffffff); part_u[idx, P_ALIVE + 1] = np.uint32(mask >> 32)
    env[:, E_STRENGTH] = strength
    env[:, E_LIMIT_SCALE] = limit_scale
    env[:, E_TREMOR_ON] = (impairment == 3)
    env[:, E_TREMOR:E_TREMOR + 4] = tremor
    env[:, E_HUMAN_KP] = 0.005                                 # human_gains passed to take_step, feeding.py:48
    return env, part, variant
