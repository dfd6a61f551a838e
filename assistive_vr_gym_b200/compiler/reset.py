"""Episode initial states for ScratchItch (CPU, numpy).

Restates the random part of reference `ScratchItchEnv.reset` (`scratch_itch.py:155-162,230-256,275-287`),
`WorldCreation.create_new_world` impairment draw (`world_creation.py:66-72`), `setup_human_joints` tremor draw and
limit clamp (`world_creation.py:136-141,172`) and the IK start pose (`util.py:34-105`). PyBullet's
`calculateInverseKinematics` is replaced by our own damped-least-squares solver with the same acceptance test
(position error < 0.03 and quaternion distance < 0.03, `util.py:51`); exact RNG parity with the reference is not
attempted — parity runs import a captured post-reset state instead.

On-device batched reset is the first "next" row of SURVEY.md §8f; until then initial robot poses come from a pool of
IK solutions computed here.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from . import xform as X
from .blob import ENV_STRIDE
from .mbody import MultiBodyDesc
from .scene import CompiledScene

# AvgEnv indices (include/avg_model.h)
E_Q, E_QD, E_MTARGET = 0, 32, 64
E_STRENGTH, E_LIMIT_SCALE, E_HUMAN_KP, E_TREMOR_ON = 96, 97, 98, 99
E_TREMOR, E_TARGET_H, E_TARGET_ON_ARM, E_LIMB_FRAME, E_EBODY = 100, 110, 120, 123, 124
E_ITERATION, E_TASK_SUCCESS, E_PREV_CONTACT, E_VALID_POSE, E_HAS_VALID, E_TARGET_POS = 152, 153, 154, 157, 161, 162
E_EPISODE_RETURN, E_OVERFLOW = 165, 166
E_TARGET_MASK = 170
F_SHOULDER, F_ELBOW = 5, 6


def _chain_to(mb: MultiBodyDesc, link: int) -> List[int]:
    chain = []
    i = link
    while i >= 0:
        chain.append(i)
        i = mb.links[i].parent
    return chain[::-1]


def ee_pose_and_jacobian(mb: MultiBodyDesc, q: Dict[int, float], ee_link: int, joints: List[int]):
    # link frames along the chain base -> ee_link only (MultiBodyDesc.link_frames walks every link: 87 on the PR2)
    frames = {}
    pp, pq = np.asarray(mb.base_pos, float), np.asarray(mb.base_quat, float)
    for i in _chain_to(mb, ee_link):
        lk = mb.links[i]
        pp, pq = X.tf_mul(pp, pq, lk.pos, lk.quat)
        v = q.get(i, 0.0)
        if lk.jtype == "revolute":
            pq = X.quat_normalize(X.quat_mul(pq, X.quat_from_axis_angle(lk.axis, v)))
        elif lk.jtype == "prismatic":
            pp = pp + X.quat_rotate(pq, lk.axis * v)
        frames[i] = (pp, pq)
    l = mb.links[ee_link]
    p_ee, q_ee = X.tf_mul(*frames[ee_link], l.inertial_pos, l.inertial_quat)
    chain = set(frames)
    J = np.zeros((6, len(joints)))
    for k, j in enumerate(joints):
        if j not in chain:
            continue
        pj, qj = frames[j]
        a = X.quat_rotate(qj, mb.links[j].axis / np.linalg.norm(mb.links[j].axis))
        J[:3, k] = np.cross(a, p_ee - pj)
        J[3:, k] = a
    return p_ee, q_ee, J


def _rot_err(q_target, q_cur):
    dq = X.quat_mul(q_target, X.quat_conj(q_cur))
    if dq[3] < 0:
        dq = -dq
    s = np.linalg.norm(dq[:3])
    if s < 1e-12:
        return np.zeros(3)
    ang = 2.0 * np.arctan2(s, dq[3])
    return dq[:3] / s * ang


def ik_dls(mb: MultiBodyDesc, ee_link: int, joints: List[int], lower, upper, target_pos, target_quat, q0,
           iters: int = 300, damping: float = 0.05) -> Tuple[np.ndarray, float, float]:
    q = np.array(q0, dtype=np.float64)
    for _ in range(iters):
        p, r, J = ee_pose_and_jacobian(mb, dict(zip(joints, q)), ee_link, joints)
        if target_quat is None:                       # position-only IK (calculateInverseKinematics without targetOrientation)
            e = target_pos - p; J = J[:3]
            if np.linalg.norm(e) < 1e-4:
                break
        else:
            e = np.concatenate([target_pos - p, _rot_err(target_quat, r)])
            if np.linalg.norm(e[:3]) < 1e-4 and np.linalg.norm(e[3:]) < 1e-3:
                break
        dq = J.T @ np.linalg.solve(J @ J.T + damping ** 2 * np.eye(len(e)), e)
        n = np.max(np.abs(dq))
        if n > 0.3:
            dq *= 0.3 / n
        q = np.clip(q + dq, lower, upper)
    p, r, _ = ee_pose_and_jacobian(mb, dict(zip(joints, q)), ee_link, joints)
    if target_quat is None:
        return q, float(np.linalg.norm(target_pos - p)), 0.0
    dq1 = np.linalg.norm(target_quat - r); dq2 = np.linalg.norm(target_quat + r)
    return q, float(np.linalg.norm(target_pos - p)), float(min(dq1, dq2))


def build_reset_data(scene: CompiledScene, rng: np.random.RandomState, ik_pool: int = 32) -> dict:
    """Everything `sample_states` needs, as plain arrays (so it can run where the reference assets are absent).

    Robot start poses: `scratch_itch.py:251-253` + `util.py:34-57` (IK with up to 40 random restarts, accepted when
    position error and quaternion distance are < 0.03); the tool pose follows from `world_creation.py:331-337`.
    """
    robot, human = scene.multibodies[0], scene.multibodies[1]
    joints = scene.robot_arm_joints
    lower = np.array([robot.links[j].lower for j in joints])
    upper = np.array([robot.links[j].upper for j in joints])
    target_quat = X.quat_from_euler([0, np.pi / 2.0, 0])
    pool_q, pool_tool = [], []
    at = scene.attach[2][-1]
    ip, iq = X.tf_inv(at.pos, at.quat)
    while len(pool_q) < ik_pool:
        target_pos = np.array([-0.5, 0, 0.8]) + rng.uniform(-0.05, 0.05, size=3)
        best = None
        for _ in range(40):                                            # max_ik_random_restarts
            rest = rng.uniform(np.maximum(lower, -np.pi), np.minimum(upper, np.pi))
            q, ep, eq = ik_dls(robot, 8, joints, lower, upper, target_pos, target_quat, rest)
            if ep < 0.03 and eq < 0.03:                                # random_restart_threshold
                best = q
                break
            if best is None:
                best = q
        qrob = {j: float(best[k]) for k, j in enumerate(joints)}
        for j in (9, 11, 13):
            qrob[j] = 1.0                                              # gripper open position, scratch_itch.py:254
        ee_p, ee_q = robot.com_frames(qrob)[8]
        base_p, base_q = X.tf_mul(ee_p, ee_q, *scene.tool_offset)
        bp, bq = X.tf_mul(base_p, base_q, ip, iq)
        pool_q.append(np.asarray(best)); pool_tool.append(np.concatenate([bp, bq]))
    arm_qidx, arm_dof, fin_qidx, fin_dof, hum_qidx, hum_dof, hum_joint = [], [], [], [], [], [], []
    tool_qidx = -1
    for b in scene.bodies:
        if b.art == 0 and b.jtype != 2:
            if b.ref_joint in joints:
                arm_qidx.append(b.qidx); arm_dof.append(b.dof)
            else:
                fin_qidx.append(b.qidx); fin_dof.append(b.dof)
        elif b.art == 1:
            hum_qidx.append(b.qidx); hum_dof.append(b.dof); hum_joint.append(b.ref_joint)
        elif b.art == 2:
            tool_qidx = b.qidx
    hum_lower = np.array([human.links[j].lower for j in hum_joint])
    hum_upper = np.array([human.links[j].upper for j in hum_joint])
    hum_reset = np.array([scene.q_human_reset.get(j, 0.0) for j in hum_joint])
    limb = human.dims["limb_dims"]
    return dict(pool_q=np.asarray(pool_q), pool_tool=np.asarray(pool_tool), arm_qidx=np.asarray(arm_qidx),
                arm_dof=np.asarray(arm_dof), fin_qidx=np.asarray(fin_qidx), fin_dof=np.asarray(fin_dof),
                hum_qidx=np.asarray(hum_qidx), hum_dof=np.asarray(hum_dof), hum_joint=np.asarray(hum_joint),
                hum_lower=hum_lower, hum_upper=hum_upper, hum_reset=hum_reset, tool_qidx=np.asarray(tool_qidx),
                limb_dims=np.array([limb[9], limb[11]]), human_control=np.asarray(int(scene.human_control)))


def sample_states(reset_data: List[dict], n: int, rng: np.random.RandomState, genders: np.ndarray | None = None):
    """Post-reset env records for n environments. reset_data[v] = build_reset_data of variant v (0 male, 1 female).

    The quantities drawn per environment are those of SURVEY.md App. C: gender (scratch_itch.py:156), impairment and
    its parameters (world_creation.py:67-72), tremor amplitudes (:141), start pose (a pool entry), limb
    (scratch_itch.py:278) and the point on the capsule (util.py:118,129).  Draws are vectorised over environments, so
    the stream order differs from the reference's per-episode order (exact RNG parity is out of reach anyway, App. C).
    -> (records [n, ENV_STRIDE] float32 with int slots bit-cast, variant [n] int32)
    """
    env = np.zeros((n, ENV_STRIDE), dtype=np.float32)
    env_i = env.view(np.int32)
    nv = len(reset_data)
    task = int(reset_data[0].get("task", 0))
    npg = max(nv // 2, 1)                                      # variants per gender (BedBathing: one per robot base pose)
    gender = (np.asarray(genders, dtype=np.int32) if genders is not None else rng.randint(min(nv, 2), size=n).astype(np.int32))
    impairment = rng.randint(4, size=n)                        # 0 none, 1 limits, 2 weakness, 3 tremor
    if task == 1:
        impairment[:] = 0                                      # human_impairment='none', bed_bathing.py:198
    limit_scale = np.where(impairment == 1, rng.uniform(0.5, 1.0, size=n), 1.0)
    strength = np.where(impairment == 2, rng.uniform(0.25, 1.0, size=n), 1.0)
    tremor = rng.uniform(np.deg2rad(-10), np.deg2rad(10), size=(n, 10)) * (impairment == 3)[:, None]
    limb = rng.randint(2, size=n)
    u_len = rng.uniform(0.0, 1.0, size=n)
    theta = rng.uniform(0, 2 * np.pi, size=n)
    pool_pick = rng.randint(1 << 30, size=n)
    variant = (gender * npg + pool_pick % npg).astype(np.int32)
    for v in range(nv):
        idx = np.nonzero(variant == v)[0]
        if idx.size == 0:
            continue
        rd = reset_data[v]
        k = (pool_pick[idx] // npg) % len(rd["pool_q"])
        qa = rd["pool_q"][k]
        fin_open = float(rd.get("fin_open", 1.0))
        ls = limit_scale[idx][:, None]
        qh = np.clip(rd["hum_reset"][None, :], rd["hum_lower"][None, :] * ls, rd["hum_upper"][None, :] * ls)   # world_creation.py:172
        length = rd["limb_dims"][limb[idx], 0]; radius = rd["limb_dims"][limb[idx], 1]
        rl = radius + u_len[idx] * (length - radius)            # uniform(radius, length), util.py:118
        env[np.ix_(idx, E_Q + rd["arm_qidx"])] = qa
        env[np.ix_(idx, E_MTARGET + rd["arm_dof"])] = qa
        env[np.ix_(idx, E_Q + rd["fin_qidx"])] = fin_open
        env[np.ix_(idx, E_MTARGET + rd["fin_dof"])] = fin_open
        env[np.ix_(idx, E_Q + rd["hum_qidx"])] = qh
        env[np.ix_(idx, E_MTARGET + rd["hum_dof"])] = qh
        tq = int(rd["tool_qidx"])
        env[idx, E_Q + tq:E_Q + tq + 7] = rd["pool_tool"][k]
        active = bool(rd["human_control"]) | (impairment[idx] == 3)
        env[idx, E_HUMAN_KP] = np.where(active, 0.05, 0.01)     # scratch_itch.py:45 / :231
        env[np.ix_(idx, E_TARGET_H + rd["hum_joint"] - 4)] = qh   # scratch_itch.py:235
        if task == 1:
            env_i[np.ix_(idx, E_TARGET_MASK + np.arange(5))] = target_mask_words(int(rd["n_target"]))[None, :]   # bed_bathing.py:369-379
            continue
        th = theta[idx]
        env[idx, E_TARGET_ON_ARM + 0] = -radius * np.sin(th)
        env[idx, E_TARGET_ON_ARM + 1] = -radius * np.cos(th)
        env[idx, E_TARGET_ON_ARM + 2] = -rl
    env[:, E_STRENGTH] = strength
    env[:, E_LIMIT_SCALE] = limit_scale
    env[:, E_TREMOR_ON] = (impairment == 3)
    env[:, E_TREMOR:E_TREMOR + 10] = tremor
    env_i[:, E_LIMB_FRAME] = F_SHOULDER if task == 1 else np.where(limb == 0, F_SHOULDER, F_ELBOW)
    return env, variant


def target_mask_words(n_target: int) -> np.ndarray:
    """AVG_E_TARGET_MASK words with the first n_target bits set (every wiping target still alive)."""
    w = np.zeros(5, dtype=np.uint32)
    for t in range(n_target):
        w[t >> 5] |= np.uint32(1 << (t & 31))
    return w.view(np.int32)


# Start pose by inverse kinematics, per (task, robot): EE link target of the reference's reset() -- centre of the +-0.05 box and
# orientation (scratch_itch.py:243-244 PR2, :251-252 Jaco).  BedBathing's start target is fixed, so its pool entry is exact.
IK_START_TARGET = {("scratch_itch", "jaco"): ([-0.5, 0.0, 0.8], [0.0, np.pi / 2.0, 0.0]),
                   ("scratch_itch", "pr2"): ([-0.55, 0.0, 0.8], [0.0, 0.0, 0.0])}


def device_ik_setup(blob: bytes, task: str, robot: str) -> dict | None:
    """What the device IK needs beyond the model: the EE link COM frame relative to its dynamic body, recovered from the
    blob's weld-parent frame (= EE COM frame o tool offset, world_creation.py:331-337,356) and the task's tool offset."""
    if (task, robot) not in IK_START_TARGET:
        return None
    from .blob import read_blob
    from .scene import TOOL_SETUP
    fr = read_blob(blob)["frames"][2]                                   # AVG_F_WELD_PARENT
    _, tool_pos, tool_euler = TOOL_SETUP[(task, robot)]
    ip, iq = X.tf_inv(np.asarray(tool_pos, float), X.quat_from_euler(tool_euler))
    ee_p, ee_q = X.tf_mul(np.asarray(fr["pos"], float), np.asarray(fr["quat"], float), ip, iq)
    pos, euler = IK_START_TARGET[(task, robot)]
    return dict(ee_body=int(fr["body"]), ee_pos=ee_p, ee_quat=ee_q, target_pos=np.asarray(pos, float),
                target_quat=X.quat_from_euler(euler), range=0.05)


# ---- device reset (include/avg_model.h AvgResetTable, csrc avg_reset_kernel) ------------------------------------------
RESET_POOL = 64
RESET_TABLE_DT = np.dtype([
    ("n_pool", "<i4"), ("n_arm", "<i4"), ("n_fin", "<i4"), ("n_hum", "<i4"), ("tool_qidx", "<i4"), ("human_control", "<i4"), ("task", "<i4"), ("n_target", "<i4"),
    ("pool_q", "<f4", (RESET_POOL, 8)), ("pool_tool", "<f4", (RESET_POOL, 8)),
    ("arm_qidx", "<i4", 8), ("arm_dof", "<i4", 8), ("fin_qidx", "<i4", 8), ("fin_dof", "<i4", 8),
    ("hum_qidx", "<i4", 8), ("hum_dof", "<i4", 8), ("hum_joint", "<i4", 8),
    ("hum_lower", "<f4", 8), ("hum_upper", "<f4", 8), ("hum_reset", "<f4", 8), ("limb_dims", "<f4", (2, 2)),
    ("fin_open", "<f4"), ("ik_enabled", "<i4"), ("ik_ee_body", "<i4"), ("ik_range", "<f4"),
    ("ik_target", "<f4", 8), ("ik_ee_frame", "<f4", 8),
    ("hum_slot", "<i4", 8), ("fin_q", "<f4", 8), ("n_particle", "<i4"), ("has_bowl", "<i4"), ("head_mask", "<u4"), ("ik_tol", "<f4"),
    ("bowl_center", "<f4", 4), ("bowl_quat", "<f4", 4), ("grid", "<f4", (64, 4)),
    ("new_mode", "<i4"), ("hum_jitter", "<f4"), ("new_min_dist", "<f4"), ("pad_new", "<i4"),
])


def reset_table_bytes(rd: dict, ik: dict | None = None) -> bytes:
    """One variant's `build_reset_data` as the AvgResetTable the device sampler reads.  `ik` (device_ik_setup) switches the
    start pose from the pool to the on-device IK."""
    t = np.zeros(1, dtype=RESET_TABLE_DT)[0]
    if ik is not None:
        t["ik_enabled"] = 1; t["ik_ee_body"] = int(ik["ee_body"]); t["ik_range"] = float(ik["range"])
        t["ik_target"][:3] = ik["target_pos"]; t["ik_target"][3:7] = ik["target_quat"]
        t["ik_ee_frame"][:3] = ik["ee_pos"]; t["ik_ee_frame"][3:7] = ik["ee_quat"]
    n_pool = min(len(rd["pool_q"]), RESET_POOL)
    t["n_pool"] = n_pool; t["n_arm"] = len(rd["arm_qidx"]); t["n_fin"] = len(rd["fin_qidx"]); t["n_hum"] = len(rd["hum_qidx"])
    t["tool_qidx"] = int(rd["tool_qidx"]); t["human_control"] = int(rd["human_control"])
    t["pool_q"][:n_pool, :rd["pool_q"].shape[1]] = rd["pool_q"][:n_pool]
    t["pool_tool"][:n_pool, :7] = rd["pool_tool"][:n_pool]
    for k in ("arm_qidx", "arm_dof", "fin_qidx", "fin_dof", "hum_qidx", "hum_dof", "hum_joint", "hum_lower", "hum_upper", "hum_reset"):
        t[k][:len(rd[k])] = rd[k]
    t["limb_dims"] = rd["limb_dims"]
    t["task"] = int(rd.get("task", 0)); t["n_target"] = int(rd.get("n_target", 0)); t["fin_open"] = float(rd.get("fin_open", 1.0))
    fd = int(t["task"]) >= 2
    hj = np.asarray(rd["hum_joint"], dtype=np.int64)
    t["hum_slot"][:len(hj)] = hj - (24 if fd else 4)
    t["fin_q"][:len(rd["fin_qidx"])] = rd["fin_q"] if "fin_q" in rd else float(rd.get("fin_open", 1.0))
    t["ik_tol"] = float(rd.get("ik_tol", 0.03))
    if int(rd.get("new_mode", 0)):                             # <Task><Robot>New-v0
        t["new_mode"] = 1; t["hum_jitter"] = float(rd.get("hum_jitter", 0.0)); t["new_min_dist"] = float(rd.get("new_min_dist", 0.01))
        if "frozen_mask" in rd:                                # BedBathing New: the arm is static during play (bed_bathing.py:271)
            t["head_mask"] = int(rd["frozen_mask"])
    if fd:                                                     # Feeding / Drinking (compiler/reset_fd.py build_reset_data_fd)
        npart = int(rd["n_particle"])
        t["n_particle"] = npart; t["has_bowl"] = int(rd["has_bowl"]); t["head_mask"] = int(rd["head_mask"])
        t["bowl_center"][:3] = rd["bowl_center"]; t["bowl_quat"] = rd["bowl_quat"]
        t["grid"][:npart, :3] = rd["grid"]
        if ik is None:                                         # these tasks always solve the start pose on the device
            t["ik_enabled"] = 1; t["ik_ee_body"] = int(rd["ik_ee_body"]); t["ik_range"] = 0.05
            t["ik_target"][:3] = rd["ik_center"]; t["ik_target"][3:7] = rd["ik_quat"]
            t["ik_ee_frame"][:3] = rd["ik_ee_pos"]; t["ik_ee_frame"][3:7] = rd["ik_ee_quat"]
    return t.tobytes()


def _mix(h: np.ndarray) -> np.ndarray:
    h = h ^ (h >> np.uint32(16)); h = h * np.uint32(0x85ebca6b); h = h ^ (h >> np.uint32(13)); h = h * np.uint32(0xc2b2ae35)
    return h ^ (h >> np.uint32(16))


def reset_u32(seed: int, env: np.ndarray, episode: np.ndarray, k: int) -> np.ndarray:
    """AVG_RNG_MIX chain of include/avg_model.h (draw k of episode `episode` of environment `env`)."""
    with np.errstate(over="ignore"):
        h = np.full(env.shape, seed & 0xffffffff, dtype=np.uint32)
        h = _mix(h) ^ (env.astype(np.uint32) * np.uint32(0x9e3779b9))
        h = _mix(h) ^ (episode.astype(np.uint32) * np.uint32(0x7f4a7c15))
        h = _mix(h) ^ np.uint32((k * 0x94d049bb) & 0xffffffff)
        return _mix(h)


def reset_u01(seed, env, episode, k) -> np.ndarray:
    return (reset_u32(seed, env, episode, k) >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def sample_states_hashed(reset_data: List[dict], n: int, seed: int, episode: np.ndarray):
    """Host mirror of avg_reset_kernel: the records the device writes for environments 0..n-1 starting their
    `episode`-th episode under `seed` (float32 arithmetic in the kernel's order)."""
    f32 = np.float32
    env_idx = np.arange(n)
    env = np.zeros((n, ENV_STRIDE), dtype=np.float32)
    env_i = env.view(np.int32)
    nv = len(reset_data)
    task = int(reset_data[0].get("task", 0))
    npg = max(nv // 2, 1)
    gender = (reset_u32(seed, env_idx, episode, 0) % np.uint32(min(nv, 2))).astype(np.int32)
    impairment = (reset_u32(seed, env_idx, episode, 1) & np.uint32(3)).astype(np.int32)
    if task == 1:
        impairment[:] = 0
    limit_scale = np.where(impairment == 1, f32(0.5) + f32(0.5) * reset_u01(seed, env_idx, episode, 2), f32(1.0)).astype(f32)
    strength = np.where(impairment == 2, f32(0.25) + f32(0.75) * reset_u01(seed, env_idx, episode, 3), f32(1.0)).astype(f32)
    deg10 = f32(0.17453292519943295)
    tremor = np.stack([(f32(2.0) * reset_u01(seed, env_idx, episode, 4 + j) - f32(1.0)) * deg10 for j in range(10)], axis=1)
    tremor = np.where((impairment == 3)[:, None], tremor, f32(0.0)).astype(f32)
    limb = (reset_u32(seed, env_idx, episode, 14) & np.uint32(1)).astype(np.int32)
    u_len = reset_u01(seed, env_idx, episode, 15)
    theta = f32(6.283185307179586) * reset_u01(seed, env_idx, episode, 16)
    pick = reset_u32(seed, env_idx, episode, 17)
    variant = (gender * npg + (pick % np.uint32(npg)).astype(np.int32)).astype(np.int32)
    for v in range(nv):
        idx = np.nonzero(variant == v)[0]
        if idx.size == 0:
            continue
        rd = reset_data[v]
        n_pool = min(len(rd["pool_q"]), RESET_POOL)
        k = ((pick[idx] // np.uint32(npg)) % np.uint32(n_pool)).astype(np.int64)
        fin_open = f32(rd.get("fin_open", 1.0))
        qa = rd["pool_q"][k].astype(f32)
        ls = limit_scale[idx][:, None]
        qh = np.minimum(np.maximum(rd["hum_reset"].astype(f32)[None, :], rd["hum_lower"].astype(f32)[None, :] * ls), rd["hum_upper"].astype(f32)[None, :] * ls)
        length = rd["limb_dims"].astype(f32)[limb[idx], 0]; radius = rd["limb_dims"].astype(f32)[limb[idx], 1]
        rl = radius + u_len[idx] * (length - radius)
        env[np.ix_(idx, E_Q + rd["arm_qidx"])] = qa
        env[np.ix_(idx, E_MTARGET + rd["arm_dof"])] = qa
        env[np.ix_(idx, E_Q + rd["fin_qidx"])] = fin_open
        env[np.ix_(idx, E_MTARGET + rd["fin_dof"])] = fin_open
        env[np.ix_(idx, E_Q + rd["hum_qidx"])] = qh
        env[np.ix_(idx, E_MTARGET + rd["hum_dof"])] = qh
        tq = int(rd["tool_qidx"])
        env[idx, E_Q + tq:E_Q + tq + 7] = rd["pool_tool"][k].astype(f32)
        active = bool(rd["human_control"]) | (impairment[idx] == 3)
        env[idx, E_HUMAN_KP] = np.where(active, f32(0.05), f32(0.01))
        env[np.ix_(idx, E_TARGET_H + rd["hum_joint"] - 4)] = qh
        if task == 1:
            env_i[np.ix_(idx, E_TARGET_MASK + np.arange(5))] = target_mask_words(int(rd["n_target"]))[None, :]
            continue
        th = theta[idx]
        env[idx, E_TARGET_ON_ARM + 0] = -radius * np.sin(th, dtype=f32)
        env[idx, E_TARGET_ON_ARM + 1] = -radius * np.cos(th, dtype=f32)
        env[idx, E_TARGET_ON_ARM + 2] = -rl
    env[:, E_STRENGTH] = strength
    env[:, E_LIMIT_SCALE] = limit_scale
    env[:, E_TREMOR_ON] = (impairment == 3)
    env[:, E_TREMOR:E_TREMOR + 10] = tremor
    env_i[:, E_LIMB_FRAME] = F_SHOULDER if task == 1 else np.where(limb == 0, F_SHOULDER, F_ELBOW)
    return env, variant


# =====================================================================================================================
# BedBathing: robot base placement (reference AssistiveEnv.position_robot_toc, env.py:486-585) and the reset tables
# =====================================================================================================================
def joint_limited_weighting(q, lower, upper) -> np.ndarray:
    """env.py:466-477"""
    phi, lam = 0.5, 0.05
    w = []
    for qi, l, u in zip(q, lower, upper):
        qr = 0.5 * (u - l)
        w.append(max(1.0 - np.power(phi, (qr - np.abs(qr - qi + l)) / (lam * qr) + 1), 0.001))
    return np.diag(w)


def toc_search(robot: MultiBodyDesc, joints: List[int], start_pos, start_quat, goal_points, rng: np.random.RandomState,
                    pos_offset, attempts: int = 100, random_rotation: float = 30.0, random_position: float = 0.1,
                    max_ik_iterations: int = 200, ee_link: int = 8):
    """`position_robot_toc` for a single-arm robot (env.py:486-585 as called at bed_bathing.py:325): `attempts` random
    base poses (x in [-random_position, 0], y in +-random_position, yaw in +-random_rotation deg); a pose is usable when
    IK reaches the start pose (position and orientation within 0.03, util.py:70); its score is (goals reached,
    sum of joint-limit-weighted kinematic isotropy).  PyBullet's IK is replaced by our damped-least-squares solver
    and the 5-step self-contact test of `ik_jlwki(step_sim=True)` (util.py:62-67) is not applied.
    -> (random_pos xy, yaw, start joint angles)"""
    lower = np.array([robot.links[j].lower for j in joints]); upper = np.array([robot.links[j].upper for j in joints])
    ik_lo = np.where(lower > upper, -2 * np.pi, lower); ik_hi = np.where(lower > upper, 2 * np.pi, upper)     # util.py:86-88
    lower, upper = np.where(lower > upper, -1e10, lower), np.where(lower > upper, 1e10, upper)                # world_creation.py:122-124
    best = None
    it = 0
    while it < attempts or best is None:
        it += 1
        rp = np.array([rng.uniform(-random_position, 0.0), rng.uniform(-random_position, random_position), 0.0])
        yaw = np.deg2rad(rng.uniform(-random_rotation, random_rotation))
        robot.base_pos = np.array([-0.85, -0.4, 0.0]) + np.asarray(pos_offset, float) + rp
        robot.base_quat = X.quat_from_euler([0, 0, yaw])
        reached, manip, q_start = 0, 0.0, None
        for j, (tp, tq) in enumerate([(start_pos, start_quat)] + [(g, None) for g in goal_points]):
            rest = rng.uniform(ik_lo, ik_hi)
            q, ep, eq = ik_dls(robot, ee_link, joints, ik_lo, ik_hi, np.asarray(tp, float), tq, rest, iters=max_ik_iterations)
            ok = ep < 0.03 and (tq is None or eq < 0.03)
            if ok:
                _, _, J = ee_pose_and_jacobian(robot, dict(zip(joints, q)), ee_link, joints)
                W = joint_limited_weighting(q, lower, upper)
                JWJ = J @ W @ J.T
                det = max(np.linalg.det(JWJ), 0.0)
                manip += np.power(det, 1.0 / 6.0) / (np.trace(JWJ) / 6.0)
                reached += 1
                if j == 0:
                    q_start = q
            elif j == 0:
                reached = -1
                break
        if reached > 0 and (best is None or reached > best[0] or (reached == best[0] and manip > best[1])):
            best = (reached, manip, rp[:2].copy(), yaw, q_start)
    return best[2], best[3], best[4], best[0]


toc_search_jaco = toc_search      # first user: BedBathingJaco-v0 (bed_bathing.py:325)


def build_reset_data_bed_bathing(scene: CompiledScene, q_start: np.ndarray) -> dict:
    """Reset table input of one play variant of a TOC-placed robot (= one robot base pose; BedBathing on both robots,
    ScratchItch on the PR2): arm at the IK start pose(s) (env.py:571-572; `q_start` = one pose or a pool [k, 7]),
    fingers at the task's open position (bed_bathing.py:320,327, scratch_itch.py:247), tool in the gripper
    (world_creation.py:331-337); BedBathing: every wiping target alive (bed_bathing.py:369-379)."""
    robot = scene.multibodies[0]
    joints = scene.robot_arm_joints
    spec = getattr(scene, "robot_spec", None) or dict(fingers=[9, 11, 13], ee_link=8, q_preset={})
    at = scene.attach[2][-1]
    ip, iq = X.tf_inv(at.pos, at.quat)
    pool_q = np.atleast_2d(np.asarray(q_start, dtype=np.float64))
    pool_tool = []
    for qs in pool_q:
        qrob = dict(spec["q_preset"])
        qrob.update({j: float(qs[k]) for k, j in enumerate(joints)})
        for j in spec["fingers"]:
            qrob[j] = float(scene.finger_open)
        ee_p, ee_q = robot.com_frames(qrob)[spec["ee_link"]]
        base_p, base_q = X.tf_mul(ee_p, ee_q, *scene.tool_offset)
        bp, bq = X.tf_mul(base_p, base_q, ip, iq)
        pool_tool.append(np.concatenate([bp, bq]))
    arm_qidx, arm_dof, fin_qidx, fin_dof, hum_qidx, hum_dof, hum_joint = [], [], [], [], [], [], []
    tool_qidx = -1
    for b in scene.bodies:
        if b.art == 0 and b.jtype != 2:
            if b.ref_joint in joints:
                arm_qidx.append(b.qidx); arm_dof.append(b.dof)
            else:
                fin_qidx.append(b.qidx); fin_dof.append(b.dof)
        elif b.art == 1:                                   # human-active ids: the right arm stays dynamic, starts at the settled pose
            hum_qidx.append(b.qidx); hum_dof.append(b.dof); hum_joint.append(b.ref_joint)
        elif b.art == 2:
            tool_qidx = b.qidx
    human = scene.multibodies[1]
    ii = lambda a: np.asarray(a, dtype=np.int64)
    task = int(scene.header["task"])
    if task == 0:                                          # ScratchItch: limb capsules for the target draw (scratch_itch.py:277-280)
        limb = human.dims["limb_dims"]
        return dict(pool_q=pool_q, pool_tool=np.asarray(pool_tool), arm_qidx=np.asarray(arm_qidx),
                    arm_dof=np.asarray(arm_dof), fin_qidx=np.asarray(fin_qidx), fin_dof=np.asarray(fin_dof),
                    hum_qidx=ii(hum_qidx), hum_dof=ii(hum_dof), hum_joint=ii(hum_joint),
                    hum_lower=np.array([human.links[j].lower for j in hum_joint], dtype=np.float64),
                    hum_upper=np.array([human.links[j].upper for j in hum_joint], dtype=np.float64),
                    hum_reset=np.array([scene.q_human_reset.get(j, 0.0) for j in hum_joint], dtype=np.float64),
                    tool_qidx=np.asarray(tool_qidx), limb_dims=np.array([limb[9], limb[11]]),
                    human_control=np.asarray(int(scene.human_control)), task=np.asarray(0), n_target=np.asarray(0),
                    fin_open=np.asarray(float(scene.finger_open)))
    return dict(pool_q=pool_q, pool_tool=np.asarray(pool_tool), arm_qidx=np.asarray(arm_qidx),
                arm_dof=np.asarray(arm_dof), fin_qidx=np.asarray(fin_qidx), fin_dof=np.asarray(fin_dof),
                hum_qidx=ii(hum_qidx), hum_dof=ii(hum_dof), hum_joint=ii(hum_joint),
                hum_lower=np.array([human.links[j].lower for j in hum_joint], dtype=np.float64),
                hum_upper=np.array([human.links[j].upper for j in hum_joint], dtype=np.float64),
                hum_reset=np.array([scene.q_human_reset.get(j, 0.0) for j in hum_joint], dtype=np.float64),
                tool_qidx=np.asarray(tool_qidx), limb_dims=np.zeros((2, 2)), human_control=np.asarray(int(scene.human_control)),
                task=np.asarray(1), n_target=np.asarray(scene.info["n_target"]), fin_open=np.asarray(float(scene.finger_open)))


def bed_bathing_settle_record(scene: CompiledScene) -> np.ndarray:
    """Env record at the start of the reference's settle loop (bed_bathing.py:284-290): human right arm at the preset.
    The robot stands 2.8 m away (world_creation.py:288) and cannot influence the arm; it is parked upright (joints 2, 4, 6 at pi, inside
    their limits) with the wiper in its gripper instead of the reference's all-zero pose (env.py:450-453), which violates
    the Jaco's joint limits and has its links interpenetrating by centimetres."""
    env = np.zeros(ENV_STRIDE, dtype=np.float32)
    robot = scene.multibodies[0]
    qrob = {}
    for b in scene.bodies:
        if b.jtype == 2:
            continue
        if b.art == 0:
            park = dict(zip(scene.robot_arm_joints, [0.0, np.pi, 0.0, np.pi, 0.0, np.pi, 0.0]))
            v = float(scene.finger_open) if b.ref_joint in (9, 11, 13) else park.get(b.ref_joint, 0.0)
            qrob[b.ref_joint] = v
        else:
            v = float(scene.q_human_reset.get(b.ref_joint, 0.0))
        env[E_Q + b.qidx] = v; env[E_MTARGET + b.dof] = v
    at = scene.attach[2][-1]
    ip, iq = X.tf_inv(at.pos, at.quat)
    ee_p, ee_q = robot.com_frames(qrob)[8]
    base_p, base_q = X.tf_mul(ee_p, ee_q, *scene.tool_offset)
    bp, bq = X.tf_mul(base_p, base_q, ip, iq)
    for b in scene.bodies:
        if b.art == 2:
            env[E_Q + b.qidx:E_Q + b.qidx + 7] = np.concatenate([bp, bq])
    env[E_STRENGTH] = 1.0; env[E_LIMIT_SCALE] = 1.0; env[E_HUMAN_KP] = 0.0
    env.view(np.int32)[E_LIMB_FRAME] = F_SHOULDER
    return env


# =====================================================================================================================
# `New` ids: numpy mirror of the clearance test of avg_reset_new_kernel (bounding capsules; static boxes by their face axes)
# =====================================================================================================================
def _seg_seg(p1, q1, p2, q2) -> float:
    d1, d2, r = q1 - p1, q2 - p2, p1 - p2
    a, e, f = d1 @ d1, d2 @ d2, d2 @ r
    if a <= 1e-12 and e <= 1e-12:
        return float(np.linalg.norm(r))
    if a <= 1e-12:
        sc, tc = 0.0, float(np.clip(f / e, 0, 1))
    else:
        c = d1 @ r
        if e <= 1e-12:
            tc, sc = 0.0, float(np.clip(-c / a, 0, 1))
        else:
            b = d1 @ d2
            den = a * e - b * b
            sc = float(np.clip((b * f - c * e) / den, 0, 1)) if den > 1e-12 else 0.0
            tc = (b * sc + f) / e
            if tc < 0:
                tc, sc = 0.0, float(np.clip(-c / a, 0, 1))
            elif tc > 1:
                tc, sc = 1.0, float(np.clip((b - c) / a, 0, 1))
    return float(np.linalg.norm((p1 + d1 * sc) - (p2 + d2 * tc)))


def body_poses(model: dict, rec: np.ndarray):
    """World poses of the dynamic bodies of a ModelBlob for the positions in an env record (parents precede children)."""
    poses = []
    for B in model["bodies"]:
        if int(B["jtype"]) == 2:
            q = rec[E_Q + int(B["qidx"]):E_Q + int(B["qidx"]) + 7].astype(np.float64)
            poses.append((q[:3], X.quat_normalize(q[3:7]))); continue
        pp, pq = (np.zeros(3), np.array([0.0, 0, 0, 1])) if int(B["parent"]) < 0 else poses[int(B["parent"])]
        ax = np.asarray(B["axis"], float); qv = float(rec[E_Q + int(B["qidx"])])
        jp, jq = np.asarray(B["ta_pos"], float), np.asarray(B["ta_quat"], float)
        if int(B["jtype"]) == 0:
            jq = X.quat_mul(jq, X.quat_from_axis_angle(ax, qv))
        else:
            jp = jp + X.quat_rotate(jq, ax * qv)
        p1, q1 = X.tf_mul(jp, jq, np.asarray(B["tb_pos"], float), np.asarray(B["tb_quat"], float))
        p, q = X.tf_mul(pp, pq, p1, q1)
        poses.append((p, X.quat_normalize(q)))
    return poses


def new_arm_clearance(model: dict, rec: np.ndarray) -> float:
    """Smallest bounding-capsule gap between the person's moving (arm) shapes and the shapes the reference tests in the `New`
    resampling loop (scratch_itch.py:219-223, bed_bathing.py:274-276): the rest of the person except links 3 and 6, the robot,
    the furniture.  Same rule as avg_reset_new_kernel."""
    h = model["header"]; sh = model["shapes"]; bc = model["bcap"]
    nms, ns, nb = int(h["n_mshape"]), int(h["n_shape"]), int(h["n_body"])
    poses = body_poses(model, rec)

    def capsule(si):
        k = bc[si].astype(np.float64)
        if si < nms:
            S = sh[si]; bp, bq = poses[int(S["body"])]
            sp, sq = X.tf_mul(bp, bq, np.asarray(S["pos"], float), np.asarray(S["quat"], float))
            return sp + X.quat_rotate(sq, k[0:3]), sp + X.quat_rotate(sq, k[4:7]), k[3]
        return k[0:3], k[4:7], k[3]
    gap = 1e9
    for sa in range(nms):
        if int(sh[sa]["ref_body"]) != 1 or int(sh[sa]["body"]) >= nb:
            continue
        a0, a1, ra = capsule(sa)
        for sb in range(ns):
            S = sh[sb]; rb = int(S["ref_body"])
            if int(S["type"]) == 5:
                continue
            if not ((rb == 1 and sb >= nms and int(S["ref_link"]) not in (3, 6)) or rb in (0, 3)):
                continue
            if int(S["type"]) == 2 and sb >= nms:
                R = X.quat_to_mat(np.asarray(S["quat"], float))
                l0, l1 = R.T @ (a0 - S["pos"]), R.T @ (a1 - S["pos"])
                sep = max((min(abs(l0[k]), abs(l1[k])) - float(S["half"][k])) if (l0[k] > 0) == (l1[k] > 0) else -float(S["half"][k]) for k in range(3))
                gap = min(gap, sep - ra); continue
            b0, b1, rbb = capsule(sb)
            gap = min(gap, _seg_seg(a0, a1, b0, b1) - ra - rbb)
    return gap


def new_variant_acceptance(blob: bytes, rd: dict, rng: np.random.RandomState, draws: int = 24) -> float:
    """Fraction of arm-pose draws (preset + U(-jitter, jitter) per joint, clipped to the limits) that keep the 0.01 clearance for
    the person of this model variant with the robot at its first start pose: the compile tool drops persons for whom the
    reference's resampling loop (which redraws the waist pose too, scratch_itch.py:211) would rarely terminate."""
    from .blob import read_blob
    model = read_blob(blob)
    rec = np.zeros(ENV_STRIDE, dtype=np.float64)
    rec[E_Q + np.asarray(rd["arm_qidx"])] = np.asarray(rd["pool_q"])[0][:len(rd["arm_qidx"])]
    rec[E_Q + np.asarray(rd["fin_qidx"])] = float(rd.get("fin_open", 1.0))
    tq = int(rd["tool_qidx"]); rec[E_Q + tq:E_Q + tq + 7] = np.asarray(rd["pool_tool"])[0][:7]
    ok = 0
    for _ in range(draws):
        q = np.clip(np.asarray(rd["hum_reset"]) + rng.uniform(-1, 1, len(rd["hum_reset"])) * float(rd["hum_jitter"]), rd["hum_lower"], rd["hum_upper"])
        rec[E_Q + np.asarray(rd["hum_qidx"])] = q
        ok += new_arm_clearance(model, rec) >= float(rd.get("new_min_dist", 0.01))
    return ok / draws
