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
F_SHOULDER, F_ELBOW = 5, 6


def _chain_to(mb: MultiBodyDesc, link: int) -> List[int]:
    chain = []
    i = link
    while i >= 0:
        chain.append(i)
        i = mb.links[i].parent
    return chain[::-1]


def ee_pose_and_jacobian(mb: MultiBodyDesc, q: Dict[int, float], ee_link: int, joints: List[int]):
    frames = mb.link_frames(q)
    l = mb.links[ee_link]
    p_ee, q_ee = X.tf_mul(*frames[ee_link], l.inertial_pos, l.inertial_quat)
    chain = set(_chain_to(mb, ee_link))
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
        e = np.concatenate([target_pos - p, _rot_err(target_quat, r)])
        if np.linalg.norm(e[:3]) < 1e-4 and np.linalg.norm(e[3:]) < 1e-3:
            break
        dq = J.T @ np.linalg.solve(J @ J.T + damping ** 2 * np.eye(6), e)
        n = np.max(np.abs(dq))
        if n > 0.3:
            dq *= 0.3 / n
        q = np.clip(q + dq, lower, upper)
    p, r, _ = ee_pose_and_jacobian(mb, dict(zip(joints, q)), ee_link, joints)
    dq1 = np.linalg.norm(target_quat - r); dq2 = np.linalg.norm(target_quat + r)
    return q, float(np.linalg.norm(target_pos - p)), float(min(dq1, dq2))


class ScratchItchReset:
    """Sampler of post-reset states for a list of compiled scenes (one per gender variant)."""

    def __init__(self, scenes: Dict[str, CompiledScene], seed: int = 1001, ik_pool: int = 32):
        self.scenes = scenes
        self.rng = np.random.RandomState(seed)
        self.ik_pool_size = ik_pool
        self._pool = None

    # -- robot start pose pool (scratch_itch.py:251-253, util.py:34-57) --------------------------------------
    def _build_pool(self):
        scene = next(iter(self.scenes.values()))
        robot = scene.multibodies[0]
        joints = scene.robot_arm_joints
        lower = np.array([robot.links[j].lower for j in joints])
        upper = np.array([robot.links[j].upper for j in joints])
        target_quat = X.quat_from_euler([0, np.pi / 2.0, 0])
        pool = []
        while len(pool) < self.ik_pool_size:
            target_pos = np.array([-0.5, 0, 0.8]) + self.rng.uniform(-0.05, 0.05, size=3)
            best = None
            for _ in range(40):                                            # max_ik_random_restarts
                rest = self.rng.uniform(np.maximum(lower, -np.pi), np.minimum(upper, np.pi))
                q, ep, eq = ik_dls(robot, 8, joints, lower, upper, target_pos, target_quat, rest)
                if ep < 0.03 and eq < 0.03:                                # random_restart_threshold
                    best = q
                    break
                if best is None:
                    best = q
            pool.append(np.asarray(best))
        self._pool = np.asarray(pool)

    def sample(self, n: int, genders: List[str] | None = None) -> Tuple[np.ndarray, np.ndarray]:
        """-> (env records [n, ENV_STRIDE] float32 with int fields bit-cast, variant index [n] int32)."""
        if self._pool is None:
            self._build_pool()
        names = list(self.scenes.keys())
        rng = self.rng
        env = np.zeros((n, ENV_STRIDE), dtype=np.float32)
        env_i = env.view(np.int32)
        variant = np.zeros(n, dtype=np.int32)
        for e in range(n):
            g = genders[e] if genders is not None else names[rng.randint(len(names))]      # scratch_itch.py:156
            variant[e] = names.index(g)
            sc = self.scenes[g]
            robot, human, tool = sc.multibodies[0], sc.multibodies[1], sc.multibodies[2]
            impairment = ["none", "limits", "weakness", "tremor"][rng.randint(4)]          # world_creation.py:67
            limit_scale = rng.uniform(0.5, 1.0) if impairment == "limits" else 1.0           # :71
            strength = rng.uniform(0.25, 1.0) if impairment == "weakness" else 1.0           # :72
            tremor = rng.uniform(np.deg2rad(-10), np.deg2rad(10), size=10) if impairment == "tremor" else np.zeros(10)  # :141
            # robot start pose
            qa = self._pool[rng.randint(len(self._pool))]
            # human pose, clamped into the scaled limits (world_creation.py:172)
            qh = {}
            for j in sc.human_joints:
                l = human.links[j]
                qh[j] = float(np.clip(sc.q_human_reset.get(j, 0.0), l.lower * limit_scale, l.upper * limit_scale))
            # target on the arm (scratch_itch.py:275-281, util.py:112-132)
            limb = [9, 11][rng.randint(2)]
            length, radius = human.dims["limb_dims"][limb]
            rl = rng.uniform(radius, length)
            theta = rng.uniform(0, 2 * np.pi)
            target_on_arm = np.array([-radius * np.sin(theta), -radius * np.cos(theta), -rl])

            # ---- fill the record
            for bi, b in enumerate(sc.bodies):
                if b.art == 0 and b.jtype != 2:
                    if b.ref_joint in sc.robot_arm_joints:
                        v = qa[sc.robot_arm_joints.index(b.ref_joint)]
                    else:
                        v = 1.0                                                             # gripper open, :254
                    env[e, E_Q + b.qidx] = v
                    env[e, E_MTARGET + b.dof] = v
                elif b.art == 1:
                    env[e, E_Q + b.qidx] = qh[b.ref_joint]
                    env[e, E_MTARGET + b.dof] = qh[b.ref_joint]
            # tool placed at the end effector (world_creation.py:331-337)
            qrob = {j: float(qa[k]) for k, j in enumerate(sc.robot_arm_joints)}
            for j in (9, 11, 13):
                qrob[j] = 1.0
            ee_p, ee_q = robot.com_frames(qrob)[8]
            base_p, base_q = X.tf_mul(ee_p, ee_q, *sc.tool_offset)
            at = sc.attach[2][-1]
            ip, iq = X.tf_inv(at.pos, at.quat)
            bp, bq = X.tf_mul(base_p, base_q, ip, iq)
            tb = [b for b in sc.bodies if b.art == 2][0]
            env[e, E_Q + tb.qidx:E_Q + tb.qidx + 3] = bp
            env[e, E_Q + tb.qidx + 3:E_Q + tb.qidx + 7] = bq
            env[e, E_STRENGTH] = strength
            env[e, E_LIMIT_SCALE] = limit_scale
            active = sc.human_control or impairment == "tremor"
            env[e, E_HUMAN_KP] = 0.05 if active else 0.01          # scratch_itch.py:45 / :231
            env[e, E_TREMOR_ON] = 1.0 if impairment == "tremor" else 0.0
            env[e, E_TREMOR:E_TREMOR + 10] = tremor
            th = np.zeros(10)
            for j in sc.human_joints:
                th[j - 4] = qh[j]
            env[e, E_TARGET_H:E_TARGET_H + 10] = th                  # scratch_itch.py:235
            env[e, E_TARGET_ON_ARM:E_TARGET_ON_ARM + 3] = target_on_arm
            env_i[e, E_LIMB_FRAME] = F_SHOULDER if limb == 9 else F_ELBOW
            # task state: zeros (task_success, prev_target_contact_pos, iteration), scratch_itch.py:147-148
        return env, variant
