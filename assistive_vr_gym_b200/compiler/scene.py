"""Scene recipes: turn the reference's deterministic `reset()` set-up into a compiled model.

Restates, for ScratchItch on the Jaco, the model-building part of reference `scratch_itch.py:130-260` +
`world_creation.py:27-93,274-293,309-365` + `human_creation.py:57-301`: which bodies exist, where they stand, which
joints are frozen, which motors hold what, which collision pairs are filtered, the tool weld, gravity per body and the
solver settings. Random per-episode quantities (gender aside) live in the per-env state produced by `reset.py`.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import xform as X
from .human import create_human, human_self_collision_enabled
from .mbody import (Attached, DynBody, LinkDesc, MultiBodyDesc, ShapeDesc, reduce_bodies, link_contact_threshold,
                    SHAPE_BOX, SHAPE_CAPSULE, SHAPE_CYLINDER, SHAPE_HULL, SHAPE_PLANE, SHAPE_SPHERE,
                    JOINT_FREE, JOINT_PRISMATIC, JOINT_REVOLUTE)
from .meshes import load_mesh_hulls, prepare_hull
from .urdf import parse_urdf

I4 = np.array([0.0, 0, 0, 1])
REF_ROBOT, REF_HUMAN, REF_TOOL, REF_FURNITURE, REF_PLANE = 0, 1, 2, 3, 4

# reference config.ini
CONFIG = {
    "scratch_itch": dict(robot_forces=1.0, robot_gains=0.05, distance_weight=1.0, action_weight=0.01,
                         tool_force_weight=0.01, scratch_reward_weight=2.0, task_success_threshold=25.0),
    "human_preferences": dict(velocity_weight=0.25, force_nontarget_weight=0.01, high_forces_weight=0.05,
                              food_hit_weight=1.0, food_velocities_weight=1.0, dressing_force_weight=0.01,
                              high_pressures_weight=0.01),
}

_hull_cache: Dict[Tuple[str, Tuple[float, float, float]], list] = {}
hull_errors: Dict[str, float] = {}


def urdf_to_multibody(path: str, ref_body: int, name: str) -> MultiBodyDesc:
    """`p.loadURDF` without `URDF_USE_INERTIA_FROM_FILE`: masses and inertial frames from the file, inertia from the
    collision AABB (mbody.link_aabb_inertia), one hull per mesh piece."""
    u = parse_urdf(path)

    def shapes_of(link, ref_index) -> List[ShapeDesc]:
        out = []
        for c in link.collisions:
            g = c.geom
            if g.kind == "box":
                if np.all(g.size <= 0):
                    continue                       # jaco end effector: <box size="0 0 0"/>
                out.append(ShapeDesc(SHAPE_BOX, c.pos, c.quat, half=g.size / 2.0, friction=link.lateral_friction, ref_link=ref_index))
            elif g.kind == "sphere":
                out.append(ShapeDesc(SHAPE_SPHERE, c.pos, c.quat, radius=g.radius, friction=link.lateral_friction, ref_link=ref_index))
            elif g.kind == "cylinder":
                out.append(ShapeDesc(SHAPE_CYLINDER, c.pos, c.quat, radius=g.radius, half=np.array([0, 0, g.length / 2.0]),
                                     friction=link.lateral_friction, ref_link=ref_index))
            elif g.kind == "capsule":
                out.append(ShapeDesc(SHAPE_CAPSULE, c.pos, c.quat, radius=g.radius, half=np.array([0, 0, g.length / 2.0]),
                                     friction=link.lateral_friction, ref_link=ref_index))
            elif g.kind == "mesh":
                key = (g.filename, tuple(g.scale))
                if key not in _hull_cache:
                    pieces = []
                    for grp in load_mesh_hulls(g.filename):
                        v, pl, err = prepare_hull(grp * g.scale)
                        hull_errors[os.path.basename(g.filename)] = max(hull_errors.get(os.path.basename(g.filename), 0.0), err)
                        pieces.append((v, pl))
                    _hull_cache[key] = pieces
                for v, pl in _hull_cache[key]:
                    out.append(ShapeDesc(SHAPE_HULL, c.pos, c.quat, verts=v, planes=pl, friction=link.lateral_friction, ref_link=ref_index))
        return out

    bl = u.links[u.base]
    base = LinkDesc(ref_index=-1, parent=-1, jtype="base", pos=np.zeros(3), quat=I4.copy(), axis=np.zeros(3),
                    mass=bl.mass, inertial_pos=bl.inertial_pos, inertial_quat=bl.inertial_quat,
                    shapes=shapes_of(bl, -1), name=bl.name)
    links = []
    for i, j in enumerate(u.order):
        l = u.links[j.child]
        jt = {"fixed": "fixed", "revolute": "revolute", "continuous": "revolute", "prismatic": "prismatic"}[j.jtype]
        enforced = j.jtype in ("revolute", "prismatic") and j.lower <= j.upper
        links.append(LinkDesc(ref_index=i, parent=u.parent_index(i), jtype=jt, pos=j.pos, quat=j.quat, axis=j.axis,
                              mass=l.mass, inertial_pos=l.inertial_pos, inertial_quat=l.inertial_quat,
                              shapes=shapes_of(l, i), lower=j.lower, upper=j.upper, limit_enforced=enforced, name=l.name))
    return MultiBodyDesc(name=name, ref_body=ref_body, base=base, links=links)


@dataclass
class CompiledShape:
    desc: ShapeDesc
    body: int                  # dyn body index or -1 (static)
    pos: np.ndarray            # in body frame (world if static)
    quat: np.ndarray
    ref_body: int
    ref_link: int
    thr: float
    margin: float
    mb_index: int              # which MultiBodyDesc it came from (for the filter rules)


@dataclass
class CompiledScene:
    task: str
    robot_type: str
    gender: str
    human_control: bool
    multibodies: List[MultiBodyDesc]
    bodies: List[DynBody]
    attach: List[Dict[int, Attached]]      # per multibody: link -> attachment
    shapes: List[CompiledShape]
    n_mshape: int
    pairs: np.ndarray                      # [n_pair, 2] int
    frames: List[Tuple[int, np.ndarray, np.ndarray]]
    dofs: List[dict]
    header: dict
    robot_arm_joints: List[int]
    human_joints: List[int]
    q_human_reset: Dict[int, float]
    tool_offset: Tuple[np.ndarray, np.ndarray]
    info: dict = field(default_factory=dict)


def _safe_margin(half_min: float) -> float:
    """btConvexInternalShape::setSafeMargin: margin = min(0.04, 0.1 * smallest half extent)."""
    return min(0.04, 0.1 * half_min)


def _shape_margin(s: ShapeDesc) -> float:
    if s.kind in (SHAPE_SPHERE, SHAPE_CAPSULE):
        return s.radius
    if s.kind == SHAPE_BOX:
        return _safe_margin(float(np.min(s.half)))
    if s.kind == SHAPE_CYLINDER:
        return _safe_margin(min(s.radius, s.half[2]))
    if s.kind == SHAPE_HULL:
        return 0.001                   # pybullet URDF hull margin
    return 0.0


def _world_aabb(s: ShapeDesc, pos, quat):
    R = X.quat_to_mat(quat)
    lo, hi = s.local_aabb()
    c = 0.5 * (lo + hi); h = 0.5 * (hi - lo)
    return R @ c + pos, np.abs(R) @ h


def build_scratch_itch(assets_dir: str, robot_type: str = "jaco", gender: str = "male", human_control: bool = False,
                       verbose: bool = False) -> CompiledScene:
    if robot_type != "jaco":
        raise NotImplementedError("round 1 compiles the Jaco recipe only (SURVEY.md §7 step 8 lists the others as next)")
    cfg = CONFIG["scratch_itch"]
    # -- bodies as the reference creates them ---------------------------------------------------------------
    robot = urdf_to_multibody(os.path.join(assets_dir, "jaco", "j2s7s300_gym.urdf"), REF_ROBOT, "jaco")
    robot.base_pos = np.array([-0.35, -0.3, 0.36])                       # scratch_itch.py:168
    robot.base_quat = np.array([0.0, 0.0, -0.7071067811865475, 0.7071067811865476])
    robot.fixed_base = True
    h2m = 0.6 if gender == "male" else 0.54                              # scratch_itch.py:161
    human = create_human(assets_dir, gender, h2m, limit_scale=1.0, static_base=True, new=False)
    human.base_pos = np.array([0, 0.03, 0.89 - 0.23725 if gender == "male" else 0.86 - 0.225])   # scratch_itch.py:232
    human.gravity = np.array([0.0, 0.0, -1.0])                           # scratch_itch.py:260
    tool = urdf_to_multibody(os.path.join(assets_dir, "scratcher", "tool_scratch.urdf"), REF_TOOL, "scratcher")
    tool.fixed_base = False
    chair = urdf_to_multibody(os.path.join(assets_dir, "wheelchair", "wheelchair.urdf"), REF_FURNITURE, "wheelchair")
    chair.base_pos = np.array([0.0, 0.09, -0.01])                        # world_creation.py:49
    chair.base_quat = X.quat_from_euler([np.pi / 2.0, 0, -np.pi / 2.0 - 0.05])
    plane = MultiBodyDesc(name="plane", ref_body=REF_PLANE,
                          base=LinkDesc(-1, -1, "base", np.zeros(3), I4.copy(), np.zeros(3), 0.0, np.zeros(3), I4.copy(),
                                        shapes=[ShapeDesc(SHAPE_PLANE, np.zeros(3), I4.copy(), friction=1.0, ref_link=-1)]),
                          links=[])
    mbs = [robot, human, tool, chair, plane]

    # -- reset-time joint presets and frozen joints ----------------------------------------------------------
    deg = np.deg2rad
    q_human = {7: deg(30), 10: deg(-90), 20: deg(-90), 28: deg(-90), 31: deg(80), 35: deg(-90), 38: deg(80)}  # :230
    # enforce_joint_limits after the preset (world_creation.py:172, limit_scale = 1 for the baked static pose)
    for l in human.links:
        if l.jtype == "revolute":
            q_human[l.ref_index] = float(np.clip(q_human.get(l.ref_index, 0.0), l.lower, l.upper))
    controllable = list(range(4, 14))                                     # scratch_itch.py:197
    frozen_h = {l.ref_index for l in human.links if l.ref_index not in controllable}   # world_creation.py:157-161
    robot_arm = [1, 2, 3, 4, 5, 6, 7]                                     # world_creation.py:283
    fingers = [9, 11, 13]                                                 # world_creation.py:320

    bodies: List[DynBody] = []
    attach: List[Dict[int, Attached]] = []
    for k, mb in enumerate(mbs):
        frozen = frozen_h if mb is human else set()
        b, a = reduce_bodies(mb, k, q_human if mb is human else {}, frozen, len(bodies))
        bodies.extend(b); attach.append(a)
    n_body = len(bodies)
    assert n_body <= 32

    # -- dofs -----------------------------------------------------------------------------------------------
    dofs: List[dict] = []
    qidx = 0
    for bi, b in enumerate(bodies):
        if b.jtype == JOINT_FREE:
            continue
        b.dof = len(dofs); b.qidx = qidx; qidx += 1
        d = dict(body=bi, flags=0, lower=b.lower, upper=b.upper, rep_lower=b.lower, rep_upper=b.upper,
                 kp=0.0, kd=1.0, max_force=0.0, action=-1, human_slot=-1, init_target=0.0)
        if b.limit_enforced:
            d["flags"] |= 1
        if b.art == 0:
            if b.ref_joint in robot_arm:
                d.update(kp=cfg["robot_gains"], max_force=cfg["robot_forces"], action=robot_arm.index(b.ref_joint))
                d["flags"] |= 2
            elif b.ref_joint in fingers:
                d.update(kp=0.05, max_force=500.0, init_target=1.0)      # world_creation.py:328, scratch_itch.py:254
                d["flags"] |= 2
        elif b.art == 1:
            slot = controllable.index(b.ref_joint)
            d.update(kp=0.01, max_force=1.0, human_slot=slot)            # scratch_itch.py:231 reactive hold
            d["flags"] |= 2 | 4 | 8
            if human_control:
                d["action"] = 7 + slot
        dofs.append(d)
    n_jdof = len(dofs)
    n_free = 0
    for bi, b in enumerate(bodies):
        if b.jtype == JOINT_FREE:
            b.dof = len(dofs); b.qidx = qidx; qidx += 7
            for _ in range(6):
                dofs.append(dict(body=bi, flags=0, lower=0.0, upper=-1.0, rep_lower=0.0, rep_upper=-1.0, kp=0.0, kd=0.0,
                                 max_force=0.0, action=-1, human_slot=-1, init_target=0.0))
            n_free += 1
    n_dof = len(dofs)
    assert n_dof <= 32 and qidx <= 32

    # -- shapes ---------------------------------------------------------------------------------------------
    shapes: List[CompiledShape] = []
    for k, mb in enumerate(mbs):
        for li in [-1] + list(range(len(mb.links))):
            link = mb.link(li)
            if not link.shapes:
                continue
            thr = link_contact_threshold(link)
            at = attach[k][li]
            for s in link.shapes:
                p, q = X.tf_mul(at.pos, at.quat, s.pos, s.quat)
                shapes.append(CompiledShape(desc=s, body=at.body, pos=p, quat=q, ref_body=mb.ref_body, ref_link=li,
                                            thr=thr, margin=_shape_margin(s), mb_index=k))
    shapes.sort(key=lambda s: 0 if s.body >= 0 else 1)       # moving shapes first (stable)
    n_mshape = sum(1 for s in shapes if s.body >= 0)

    # -- collision pairs ------------------------------------------------------------------------------------
    def link_parent(mb: MultiBodyDesc, li: int) -> int:
        return mb.links[li].parent if li >= 0 else -2

    tool_filtered_robot_links = set(range(7, 15))             # world_creation.py:352-354 (jaco)
    pairs = []
    for ia in range(n_mshape):
        a = shapes[ia]
        for ib in range(len(shapes)):
            b = shapes[ib]
            if ib == ia or (b.body >= 0 and ib < ia):
                continue                                      # unordered pairs once; moving A first
            if a.body == b.body:
                continue
            if a.mb_index == b.mb_index:
                mb = mbs[a.mb_index]
                if a.ref_link == b.ref_link:
                    continue
                if link_parent(mb, a.ref_link) == b.ref_link or link_parent(mb, b.ref_link) == a.ref_link:
                    continue                                  # Bullet never collides a link with its parent
                if mb is robot:
                    pass                                      # URDF_USE_SELF_COLLISION, world_creation.py:282
                elif mb is human:
                    if not human_self_collision_enabled(a.ref_link, b.ref_link):
                        continue
                else:
                    continue                                  # tool / furniture: no self collision flag
            else:
                ms = {a.mb_index: a, b.mb_index: b}
                if 0 in ms and 2 in ms and ms[0].ref_link in tool_filtered_robot_links:
                    continue
            pairs.append((ia, ib))
    pairs = np.asarray(pairs, dtype=np.int32)

    # -- frames of interest ----------------------------------------------------------------------------------
    def com_frame(k: int, li: int):
        at = attach[k][li]
        link = mbs[k].link(li)
        p, q = X.tf_mul(at.pos, at.quat, link.inertial_pos, link.inertial_quat)
        return (at.body, p, q)

    tool_pos_offset = np.array([0.0, 0.0, 0.02])              # scratch_itch.py:255
    tool_orient_offset = X.quat_from_euler([0, -np.pi / 2.0, 0])
    ee = com_frame(0, 8)
    weld_parent = (ee[0],) + X.tf_mul(ee[1], ee[2], tool_pos_offset, tool_orient_offset)
    frames = [
        com_frame(2, 1),            # AVG_F_TOOL_TIP
        com_frame(2, -1),           # AVG_F_TOOL_BASE
        weld_parent,                # AVG_F_WELD_PARENT
        com_frame(0, 0),            # AVG_F_TORSO (robot link 0 COM, static)
        com_frame(1, 3),            # AVG_F_CHEST
        com_frame(1, 9), com_frame(1, 11), com_frame(1, 13),
    ]

    hp = CONFIG["human_preferences"]
    task_f = np.zeros(32, dtype=np.float32)
    task_f[:15] = [cfg["distance_weight"], cfg["action_weight"], cfg["tool_force_weight"], cfg["scratch_reward_weight"],
                   cfg["task_success_threshold"], hp["velocity_weight"], hp["force_nontarget_weight"],
                   hp["high_forces_weight"], hp["food_hit_weight"], hp["food_velocities_weight"],
                   0.025, 0.01, 10.0, 0.05, 1.0]
    task_f[16:19] = [-0.3, -0.1, 0.7]       # reference point of the device spatial algebra (float32 conditioning)
    header = dict(task=0, n_body=n_body, n_ebody=0, n_dof=n_dof, n_jdof=n_jdof, n_free=n_free,
                  substeps=5, solver_iters=50,                             # env.py:16, scratch_itch.py:258
                  n_action_robot=7, n_action_human=10 if human_control else 0,
                  n_obs_robot=30, n_obs_human=34 if human_control else 0, human_control=int(human_control),
                  dt=0.02, erp=0.2, lin_damp=0.04, ang_damp=0.04, residual_thr=1e-7, max_vel=100.0,
                  action_scale=0.05, weld_max_force=500.0,
                  weld_body_a=weld_parent[0], weld_body_b=frames[1][0], task_f=task_f)
    scene = CompiledScene(task="scratch_itch", robot_type=robot_type, gender=gender, human_control=human_control,
                          multibodies=mbs, bodies=bodies, attach=attach, shapes=shapes, n_mshape=n_mshape, pairs=pairs,
                          frames=frames, dofs=dofs, header=header, robot_arm_joints=robot_arm,
                          human_joints=list(range(7, 14)), q_human_reset=q_human,
                          tool_offset=(tool_pos_offset, tool_orient_offset))
    scene.mlp_layers = None
    if human_control:      # env.py:67: realistic_arm_limits_model.h5, used by enforce_realistic_human_joint_limits (env.py:353-387)
        from .h5lite import load_keras_dense_stack
        scene.mlp_layers = load_keras_dense_stack(os.path.join(assets_dir, 'realistic_arm_limits_model.h5'))
    scene.info = dict(hull_errors=dict(hull_errors), n_pairs=len(pairs), n_shapes=len(shapes), n_mshape=n_mshape)
    if verbose:
        print(scene.info)
    return scene
