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
                    SHAPE_BOX, SHAPE_CAPSULE, SHAPE_CYLINDER, SHAPE_HULL, SHAPE_PLANE, SHAPE_SPHERE, SHAPE_COMPOUND,
                    JOINT_FREE, JOINT_PRISMATIC, JOINT_REVOLUTE)
from .meshes import load_mesh_hulls, prepare_hull
from .urdf import parse_urdf

I4 = np.array([0.0, 0, 0, 1])
REF_ROBOT, REF_HUMAN, REF_TOOL, REF_FURNITURE, REF_PLANE, REF_TABLE, REF_BOWL = 0, 1, 2, 3, 4, 5, 6

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


def _file_inertia(link):
    """`URDF_USE_INERTIA_FROM_FILE` (world_creation.py:187): the <inertia> tensor of the file; a tensor with products of
    inertia is diagonalised and the inertial frame turned onto its principal axes, as Bullet's URDF importer does
    [UPSTREAM-BULLET].  -> (inertial_quat, principal moments)"""
    T = np.asarray(link.inertia_file, float)
    off = np.abs(T - np.diag(np.diag(T))).max()
    if off <= 1e-12 * max(np.abs(T).max(), 1e-30):
        return link.inertial_quat, np.diag(T).copy()
    w, V = np.linalg.eigh(T)
    if np.linalg.det(V) < 0:
        V[:, 2] = -V[:, 2]
    return X.quat_normalize(X.quat_mul(link.inertial_quat, X.mat_to_quat(V))), w


def urdf_to_multibody(path: str, ref_body: int, name: str, inertia_from_file: bool = False) -> MultiBodyDesc:
    """`p.loadURDF`: masses and inertial frames from the file, one hull per mesh piece; inertia from the collision AABB
    (mbody.link_aabb_inertia) unless `inertia_from_file` (the reference loads the PR2 with URDF_USE_INERTIA_FROM_FILE,
    world_creation.py:187); joint damping from <dynamics damping> (btMultibodyLink::m_jointDamping)."""
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

    def inertial(l):
        if inertia_from_file and l.mass > 0:
            q, d = _file_inertia(l)
            return q, d
        return l.inertial_quat, None

    bl = u.links[u.base]
    bq, bd = inertial(bl)
    base = LinkDesc(ref_index=-1, parent=-1, jtype="base", pos=np.zeros(3), quat=I4.copy(), axis=np.zeros(3),
                    mass=bl.mass, inertial_pos=bl.inertial_pos, inertial_quat=bq, inertia_diag=bd,
                    shapes=shapes_of(bl, -1), name=bl.name)
    links = []
    for i, j in enumerate(u.order):
        l = u.links[j.child]
        jt = {"fixed": "fixed", "revolute": "revolute", "continuous": "revolute", "prismatic": "prismatic"}[j.jtype]
        enforced = j.jtype in ("revolute", "prismatic") and j.lower <= j.upper
        lq, ld = inertial(l)
        links.append(LinkDesc(ref_index=i, parent=u.parent_index(i), jtype=jt, pos=j.pos, quat=j.quat, axis=j.axis,
                              mass=l.mass, inertial_pos=l.inertial_pos, inertial_quat=lq, inertia_diag=ld,
                              shapes=shapes_of(l, i), lower=j.lower, upper=j.upper, limit_enforced=enforced, name=l.name,
                              damping=j.damping))
    return MultiBodyDesc(name=name, ref_body=ref_body, base=base, links=links)


def load_robot(assets_dir: str, robot_type: str, arm: str = "left"):
    """The robot as `WorldCreation.init_<robot>` loads it (world_creation.py:181-217, 274-293) plus the indices the task
    files hard-code for it.  -> (MultiBodyDesc, spec) with spec = arm joints driven by the action (the LEFT arm on the
    PR2: robot_arm='left' at scratch_itch.py:45 / bed_bathing.py:44), gripper joints of `set_gripper_open_position`
    (world_creation.py:309-328), tool link (`init_tool`, :332), robot links whose collisions with the tool are switched
    off (:352-354), torso link of the observation (scratch_itch.py:105), joints kept out of the dynamics, presets."""
    if robot_type == "jaco":
        robot = urdf_to_multibody(os.path.join(assets_dir, "jaco", "j2s7s300_gym.urdf"), REF_ROBOT, "jaco")
        robot.self_collision = True                                        # URDF_USE_SELF_COLLISION, world_creation.py:282
        return robot, dict(arm=[1, 2, 3, 4, 5, 6, 7], fingers=[9, 11, 13], ee_link=8, tool_filtered=set(range(7, 15)),
                           torso_link=0, frozen=set(), q_preset={})
    if robot_type == "pr2":
        robot = urdf_to_multibody(os.path.join(assets_dir, "PR2", "pr2_no_torso_lift_tall.urdf"), REF_ROBOT, "pr2",
                                  inertia_from_file=True)                  # URDF_USE_INERTIA_FROM_FILE, world_creation.py:187
        robot.self_collision = False                                       # loaded without URDF_USE_SELF_COLLISION
        left = [64, 65, 66, 68, 69, 71, 72]; right = [42, 43, 44, 46, 47, 49, 50]       # world_creation.py:188-189
        fingers = [79, 80, 81, 82]                                         # world_creation.py:311 (left gripper)
        # Joints integrated on the device: the left arm and its four finger joints.  Everything else is baked into the
        # static world at its reset pose: the base is fixed and the torso lift is a fixed joint, so the other branches
        # (casters, head, right arm, lasers) are dynamically decoupled from the left arm, carry no gravity
        # (scratch_itch.py:259, bed_bathing.py:342) and are held at zero velocity by PyBullet's default joint motors --
        # they only move if something pushes them.  Also frozen: l_gripper_motor_slider / _screw (77, 78) and
        # l_gripper_joint (83): 10 g / 10 g / 1 g links without collision geometry (documented deviation).
        if arm == "right":                                                 # Feeding / Drinking: robot_arm='right' (feeding.py:48), tool on link 54
            fingers = [57, 58, 59, 60]                                     # world_creation.py:311 (right gripper)
            moving = set(right) | set(fingers)
            frozen = {l.ref_index for l in robot.links if l.jtype in ("revolute", "prismatic") and l.ref_index not in moving}
            q_preset = dict(zip(left, [1.75, 1.25, 1.5, -0.5, 1.0, 0.0, 1.0]))           # env.py:455-457 (idle left arm)
            return robot, dict(arm=right, fingers=fingers, ee_link=54, tool_filtered=set(range(49, 64)), torso_link=15,
                               frozen=frozen, q_preset=q_preset)                           # world_creation.py:335,359
        moving = set(left) | set(fingers)
        frozen = {l.ref_index for l in robot.links if l.jtype in ("revolute", "prismatic") and l.ref_index not in moving}
        q_preset = dict(zip(right, [-1.75, 1.25, -1.5, -0.5, -1.0, 0.0, -1.0]))          # env.py:455-459 reset_robot_joints
        return robot, dict(arm=left, fingers=fingers, ee_link=76, tool_filtered=set(range(71, 86)), torso_link=15,
                           frozen=frozen, q_preset=q_preset)
    if robot_type == "sawyer":
        # world_creation.py:219-245.  No reference environment instantiates it (the task files have no Sawyer branch);
        # the loader is compiled so that the build-defined Feeding / Drinking ids of BASELINE.json can use it.
        robot = urdf_to_multibody(os.path.join(assets_dir, "sawyer", "sawyer.urdf"), REF_ROBOT, "sawyer")
        robot.self_collision = True                                        # URDF_USE_SELF_COLLISION, world_creation.py:227
        # setCollisionFilterPair(..., 0) for links 3..23 x 3..23 and 0..2 x 0..8 (world_creation.py:229-234)
        robot.self_filter = lambda a, b: (3 <= a <= 23 and 3 <= b <= 23) or (0 <= a <= 2 and 0 <= b <= 8) or (0 <= b <= 2 and 0 <= a <= 8)
        return robot, dict(arm=[3, 8, 9, 10, 11, 13, 16], fingers=[20, 22], finger_signs=[1.0, -1.0], ee_link=18,
                           tool_filtered={18, 20, 21, 22, 23}, torso_link=0, frozen={4}, q_preset={})          # :235,317,332,352; joint 4 = head pan (idle, held by its default motor)
    if robot_type == "baxter":
        # world_creation.py:247-272 (right arm drives the tool; `useFixedBase` only: inertia from the collision shapes, no self collision)
        robot = urdf_to_multibody(os.path.join(assets_dir, "baxter", "baxter_custom.urdf"), REF_ROBOT, "baxter")
        robot.self_collision = False
        right = [12, 13, 14, 15, 16, 18, 19]; left = [34, 35, 36, 37, 38, 40, 41]
        fingers = [27, 29]                                                 # world_creation.py:313-315 (right gripper)
        moving = set(right) | set(fingers)
        frozen = {l.ref_index for l in robot.links if l.jtype in ("revolute", "prismatic") and l.ref_index not in moving}
        q_preset = dict(zip(left, [0.75, 1.0, 0.5, 0.5, 1.0, -0.5, 0.0]))   # env.py:460-462 reset_robot_joints (idle left arm)
        return robot, dict(arm=right, fingers=fingers, finger_signs=[1.0, -1.0], ee_link=25, tool_filtered={25, 27, 28, 29, 30},
                           torso_link=0, frozen=frozen, q_preset=q_preset)                                     # :334,360
    raise NotImplementedError(f"robot {robot_type!r}: the reference loads PR2, Jaco, Sawyer and Baxter (world_creation.py:181-293)")


TOOL_SETUP = {   # (task, robot) -> gripper open position, tool pos_offset, tool orient_offset (euler)
    ("scratch_itch", "jaco"): (1.0, [0.0, 0.0, 0.02], [0, -np.pi / 2.0, 0]),        # scratch_itch.py:254-255
    ("scratch_itch", "pr2"): (0.25, [0.0, 0.0, 0.0], [0, 0, 0]),                    # scratch_itch.py:247-248
    ("bed_bathing", "jaco"): (1.1, [-0.01, 0.0, 0.03], [0, -np.pi / 2.0, 0]),       # bed_bathing.py:327-328
    ("bed_bathing", "pr2"): (0.2, [0.0, 0.0, 0.0], [0, 0, 0]),                      # bed_bathing.py:320-321
}
TOC_POS_OFFSET = {("scratch_itch", "pr2"): [0.1, 0.0, 0.0], ("bed_bathing", "pr2"): [0.0, 0.0, 0.0],
                  ("bed_bathing", "jaco"): [0.1, 0.55, 0.6]}                         # scratch_itch.py:245, bed_bathing.py:318,325


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
    children: Optional[list] = None   # SHAPE_COMPOUND: CompiledShape per convex piece (poses in the owning body's frame)
    parent: int = -1                  # children: index of their compound in the top-level shape list


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
    cshapes: List[CompiledShape] = field(default_factory=list)   # compound children, shape table indices len(shapes) + k
    opairs: Optional[np.ndarray] = None    # pair table with the compounds expanded into their children (what the oracle walks)


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


def _assemble(mbs: List[MultiBodyDesc], q_presets: Dict[int, Dict[int, float]], frozen_sets: Dict[int, set], setup_dof,
              cross_pair_ok, robot: MultiBodyDesc, human: MultiBodyDesc):
    """Common part of every scene recipe: multibodies -> dynamic bodies (depth-first, joint bodies first), velocity dofs,
    compiled shapes (moving first) and the filtered collision-pair table.

    q_presets / frozen_sets: per multibody index, joint presets and the joints the reference freezes; setup_dof(body, dof)
    fills motors / action slots of a joint dof; cross_pair_ok(shape_a, shape_b) is the `setCollisionFilterPair` rule
    between different multibodies.  Pairs inside one multibody follow Bullet: never a link with its parent, robot with
    URDF_USE_SELF_COLLISION (world_creation.py:282), human by its filter matrix (human_creation.py:279-294), nothing else.
    """
    bodies: List[DynBody] = []
    attach: List[Dict[int, Attached]] = []
    for k, mb in enumerate(mbs):
        b, a = reduce_bodies(mb, k, q_presets.get(k, {}), frozen_sets.get(k, set()), len(bodies))
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
        unlimited = b.lower == 0.0 and b.upper == -1.0        # getJointInfo of a continuous joint; the action mask sees +-1e10 (world_creation.py:122-124)
        d = dict(body=bi, flags=0, lower=b.lower, upper=b.upper, rep_lower=-1e10 if unlimited else b.lower, rep_upper=1e10 if unlimited else b.upper,
                 kp=0.0, kd=1.0, max_force=0.0, action=-1, human_slot=-1, init_target=0.0, damping=float(b.damping))
        if b.limit_enforced:
            d["flags"] |= 1
        setup_dof(b, d)
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
    # A multibody with `env_static = e` (the Feeding bowl) is a static body whose pose comes from the environment record
    # (AVG_E_EBODY[e]): its shapes hang on "body" n_body + e, relative to the base frame.  Links listed in `compound_links`
    # (VHACD meshes: spoon, cup, bowl, head) become ONE top-level compound shape with their hulls as children.
    shapes: List[CompiledShape] = []
    for k, mb in enumerate(mbs):
        estatic = getattr(mb, "env_static", None)
        for li in [-1] + list(range(len(mb.links))):
            link = mb.link(li)
            if not link.shapes:
                continue
            thr = link_contact_threshold(link)
            at = attach[k][li]
            if estatic is not None:
                assert at.body < 0 and not mb.links, "env-static bodies are single fixed links"
                at = Attached(n_body + int(estatic), np.zeros(3), I4.copy())
            made = []
            for s in link.shapes:
                p, q = X.tf_mul(at.pos, at.quat, s.pos, s.quat)
                made.append(CompiledShape(desc=s, body=at.body, pos=p, quat=q, ref_body=mb.ref_body, ref_link=li,
                                          thr=thr, margin=_shape_margin(s), mb_index=k))
            if li in getattr(mb, "compound_links", ()) and len(made) > 1 and at.body >= 0:
                assert all(c.desc.kind == SHAPE_HULL for c in made) and len(made) <= 72
                cd = ShapeDesc(SHAPE_COMPOUND, np.zeros(3), I4.copy(), friction=made[0].desc.friction, ref_link=li,
                               children=[ShapeDesc(SHAPE_HULL, c.pos, c.quat, verts=c.desc.verts, planes=c.desc.planes) for c in made])
                shapes.append(CompiledShape(desc=cd, body=at.body, pos=np.zeros(3), quat=I4.copy(), ref_body=mb.ref_body, ref_link=li,
                                            thr=thr, margin=0.0, mb_index=k, children=made))
            else:
                shapes.extend(made)
    shapes.sort(key=lambda s: 0 if s.body >= 0 else 1)       # moving shapes first (stable)
    n_mshape = sum(1 for s in shapes if s.body >= 0)

    # -- collision pairs ------------------------------------------------------------------------------------
    def link_parent(mb: MultiBodyDesc, li: int) -> int:
        return mb.links[li].parent if li >= 0 else -2

    pairs = []
    for ia in range(n_mshape):
        a = shapes[ia]
        for ib in range(len(shapes)):
            b = shapes[ib]
            if ib == ia or (b.body >= 0 and ib < ia):
                continue                                      # unordered pairs once; moving A first
            if a.body == b.body:
                continue
            if a.body >= n_body and (b.body < 0 or b.body >= n_body):
                continue                                      # env-static against static: nothing moves
            if a.mb_index == b.mb_index:
                mb = mbs[a.mb_index]
                if a.ref_link == b.ref_link:
                    continue
                if link_parent(mb, a.ref_link) == b.ref_link or link_parent(mb, b.ref_link) == a.ref_link:
                    continue                                  # Bullet never collides a link with its parent
                if mb is robot:
                    if not getattr(mb, "self_collision", True):
                        continue                              # PR2: loaded without URDF_USE_SELF_COLLISION (world_creation.py:187)
                    if getattr(mb, "self_filter", None) is not None and mb.self_filter(a.ref_link, b.ref_link):
                        continue                              # Sawyer: setCollisionFilterPair(robot, robot, i, j, 0), world_creation.py:229-234
                    # Jaco: URDF_USE_SELF_COLLISION, world_creation.py:282
                elif mb is human:
                    if not human_self_collision_enabled(a.ref_link, b.ref_link):
                        continue
                else:
                    continue                                  # tool / furniture: no self collision flag
            elif not cross_pair_ok(a, b):
                continue
            pairs.append((ia, ib))
    pairs = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
    return bodies, attach, dofs, n_jdof, n_free, shapes, n_mshape, pairs


def expand_compounds(shapes: List[CompiledShape], pairs: np.ndarray):
    """Children of the compound shapes as a flat list (shape table indices len(shapes) + k, `parent` set) and the pair
    table with every compound replaced by its children -- the explicit pair list the oracle walks; the device expands
    compound candidates itself (csrc collide kernel)."""
    cshapes: List[CompiledShape] = []
    first = {}
    for i, s in enumerate(shapes):
        if s.desc.kind == SHAPE_COMPOUND:
            first[i] = len(shapes) + len(cshapes)
            for c in s.children:
                c.parent = i
                cshapes.append(c)
    if not first:
        return cshapes, pairs

    def members(i):
        return range(first[i], first[i] + len(shapes[i].children)) if i in first else (i,)

    out = [(ca, cb) for a, b in pairs for ca in members(int(a)) for cb in members(int(b))]
    return cshapes, np.asarray(out, dtype=np.int32).reshape(-1, 2)


def build_scratch_itch(assets_dir: str, robot_type: str = "jaco", gender: str = "male", human_control: bool = False,
                       base_xy_yaw: Tuple[float, float, float] = (0.0, 0.0, 0.0), verbose: bool = False,
                       new: bool = False, hipbone_to_mouth_height: Optional[float] = None,
                       waist: Tuple[float, float, float] = (0.0, 0.0, 0.0)) -> CompiledScene:
    """ScratchItch<Robot>[Human]-v0.  PR2: `base_xy_yaw` = the random_pos / yaw chosen by `position_robot_toc`
    (env.py:511-513 as called at scratch_itch.py:245).
    `new` = ScratchItch<Robot>New-v0 (`__init__.py:38-50`): the human is built with revolute waist joints
    (`human_creation.py:185-189`) held at the drawn `waist` angles (`scratch_itch.py:211`, frozen like every other
    non-controllable joint) and with every link length scaled by `hipbone_to_mouth_height` (`:158`, `human_creation.py:60-63`)."""
    cfg = CONFIG["scratch_itch"]
    # -- bodies as the reference creates them ---------------------------------------------------------------
    robot, rs = load_robot(assets_dir, robot_type)
    if robot_type == "jaco":
        robot.base_pos = np.array([-0.35, -0.3, 0.36])                   # scratch_itch.py:168
        robot.base_quat = np.array([0.0, 0.0, -0.7071067811865475, 0.7071067811865476])
    else:
        robot.base_pos = np.array([-0.85, -0.4, 0.0]) + np.asarray(TOC_POS_OFFSET[("scratch_itch", robot_type)]) \
            + np.array([base_xy_yaw[0], base_xy_yaw[1], 0.0])            # env.py:513
        robot.base_quat = X.quat_from_euler([0, 0, base_xy_yaw[2]])
    robot.fixed_base = True
    h2m = 0.6 if gender == "male" else 0.54                              # scratch_itch.py:161
    if new and hipbone_to_mouth_height is not None:
        h2m = float(hipbone_to_mouth_height)                             # scratch_itch.py:158: uniform(h2m - 0.1, h2m + 0.1)
    human = create_human(assets_dir, gender, h2m, limit_scale=1.0, static_base=True, new=new)
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
    if new:                                                              # scratch_itch.py:211: waist joints 0..2, U(-10, 10) degrees
        q_human.update({0: float(waist[0]), 1: float(waist[1]), 2: float(waist[2])})
    # enforce_joint_limits after the preset (world_creation.py:172, limit_scale = 1 for the baked static pose)
    for l in human.links:
        if l.jtype == "revolute":
            q_human[l.ref_index] = float(np.clip(q_human.get(l.ref_index, 0.0), l.lower, l.upper))
    controllable = list(range(4, 14))                                     # scratch_itch.py:197
    frozen_h = {l.ref_index for l in human.links if l.ref_index not in controllable}   # world_creation.py:157-161
    robot_arm = rs["arm"]                                                 # world_creation.py:283 / :189
    fingers = rs["fingers"]                                               # world_creation.py:311-320
    finger_open, tool_pos_offset, tool_euler = TOOL_SETUP[("scratch_itch", robot_type)]

    def setup_dof(b: DynBody, d: dict) -> None:
        if b.art == 0:
            if b.ref_joint in robot_arm:
                d.update(kp=cfg["robot_gains"], max_force=cfg["robot_forces"], action=robot_arm.index(b.ref_joint))
                d["flags"] |= 2
            elif b.ref_joint in fingers:
                d.update(kp=0.05, max_force=500.0, init_target=finger_open)   # world_creation.py:328, scratch_itch.py:254 / :247
                d["flags"] |= 2
        elif b.art == 1:
            slot = controllable.index(b.ref_joint)
            d.update(kp=0.01, max_force=1.0, human_slot=slot)            # scratch_itch.py:231 reactive hold
            d["flags"] |= 2 | 4 | 8
            if human_control:
                d["action"] = 7 + slot

    tool_filtered_robot_links = rs["tool_filtered"]           # world_creation.py:352-354

    def cross_pair_ok(a: CompiledShape, b: CompiledShape) -> bool:
        ms = {a.mb_index: a, b.mb_index: b}
        return not (0 in ms and 2 in ms and ms[0].ref_link in tool_filtered_robot_links)

    bodies, attach, dofs, n_jdof, n_free, shapes, n_mshape, pairs = _assemble(
        mbs, {0: rs["q_preset"], 1: q_human}, {0: rs["frozen"], 1: frozen_h}, setup_dof, cross_pair_ok, robot, human)
    n_body = len(bodies); n_dof = len(dofs)

    # -- frames of interest ----------------------------------------------------------------------------------
    def com_frame(k: int, li: int):
        at = attach[k][li]
        link = mbs[k].link(li)
        p, q = X.tf_mul(at.pos, at.quat, link.inertial_pos, link.inertial_quat)
        return (at.body, p, q)

    tool_pos_offset = np.asarray(tool_pos_offset, float)      # scratch_itch.py:255 / :248
    tool_orient_offset = X.quat_from_euler(tool_euler)
    ee = com_frame(0, rs["ee_link"])
    weld_parent = (ee[0],) + X.tf_mul(ee[1], ee[2], tool_pos_offset, tool_orient_offset)
    frames = [
        com_frame(2, 1),            # AVG_F_TOOL_TIP
        com_frame(2, -1),           # AVG_F_TOOL_BASE
        weld_parent,                # AVG_F_WELD_PARENT
        com_frame(0, rs["torso_link"]),   # AVG_F_TORSO (robot link 0, PR2: link 15; COM frame, static; scratch_itch.py:105)
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
    scene.finger_open = finger_open; scene.robot_spec = rs
    scene.mlp_layers = None
    if human_control:      # env.py:67: realistic_arm_limits_model.h5, used by enforce_realistic_human_joint_limits (env.py:353-387)
        from .h5lite import load_keras_dense_stack
        scene.mlp_layers = load_keras_dense_stack(os.path.join(assets_dir, 'realistic_arm_limits_model.h5'))
    scene.info = dict(hull_errors=dict(hull_errors), n_pairs=len(pairs), n_shapes=len(shapes), n_mshape=n_mshape)
    if verbose:
        print(scene.info)
    return scene


# =====================================================================================================================
# BedBathing (reference bed_bathing.py:155-358)
# =====================================================================================================================
CONFIG["bed_bathing"] = dict(robot_forces=1.0, robot_gains=0.05, distance_weight=1.0, action_weight=0.01,
                             wiping_reward_weight=5.0, task_success_threshold=0.3)          # config.ini:12-18

# Pose of the human's right arm (joints 7..13) after the reference's 100-step settle onto the mattress
# (bed_bathing.py:286-292).  Produced by OUR device kernels from the 'settle' stage of this recipe with
# tools/settle_bed_bathing.py on a B200 and committed as data; None until that has been run.
SETTLED_ARM_Q: Dict[str, Optional[List[float]]] = {"male": None, "female": None}
_SETTLE_FILE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "bed_bathing_settle.json")


def load_settled_arm_q() -> Dict[str, Optional[List[float]]]:
    """data/bed_bathing_settle.json (written by tools/settle_bed_bathing.py from a run of the CUDA kernels)."""
    import json
    if os.path.exists(_SETTLE_FILE):
        with open(_SETTLE_FILE) as f:
            d = json.load(f)
        for g in ("male", "female"):
            SETTLED_ARM_Q[g] = [float(x) for x in d[g]["arm_q"]]
    return SETTLED_ARM_Q


def _static_body(name: str, shapes: List[ShapeDesc], pos, quat, ref_body: int = REF_FURNITURE) -> MultiBodyDesc:
    """`p.createMultiBody(baseMass=0, baseCollisionShapeIndex=...)`: a massless base with collision shapes."""
    base = LinkDesc(-1, -1, "base", np.zeros(3), I4.copy(), np.zeros(3), 0.0, np.zeros(3), I4.copy(), shapes=shapes, name=name)
    mb = MultiBodyDesc(name=name, ref_body=ref_body, base=base, links=[])
    mb.base_pos = np.asarray(pos, float); mb.base_quat = np.asarray(quat, float); mb.fixed_base = True
    return mb


def _mesh_shapes(path: str, scale, friction: float, single_hull: bool = False) -> List[ShapeDesc]:
    """`p.createCollisionShape(GEOM_MESH, fileName, meshScale)`: one convex hull per `o` group of the file."""
    key = (path, tuple(float(x) for x in scale), single_hull)
    if key not in _hull_cache:
        groups = load_mesh_hulls(path)
        if single_hull:
            groups = [np.concatenate(groups, axis=0)]
        pieces = []
        for grp in groups:
            v, pl, err = prepare_hull(grp * np.asarray(scale, float))
            hull_errors[os.path.basename(path)] = max(hull_errors.get(os.path.basename(path), 0.0), err)
            pieces.append((v, pl))
        _hull_cache[key] = pieces
    return [ShapeDesc(SHAPE_HULL, np.zeros(3), I4.copy(), verts=v, planes=pl, friction=friction, ref_link=-1)
            for v, pl in _hull_cache[key]]


def capsule_points(p1, p2, radius: float, distance_between_points: float = 0.05, position_scale: float = 1.0) -> np.ndarray:
    """Reference `Util.capsule_points` (util.py:134-167): rings of points around a capsule's cylinder."""
    p1, p2 = np.asarray(p1, float), np.asarray(p2, float)
    axis = (p2 - p1) / np.linalg.norm(p2 - p1)
    m = int(np.argmax(np.abs(axis)))                         # util.py:169-177 orthogonal_vector
    y = np.zeros(3); y[(m + 1) % 3] = 1.0
    ortho = np.cross(axis, y); ortho /= np.linalg.norm(ortho)
    normal = np.cross(axis, ortho)
    sections = int(np.linalg.norm(p2 - p1) / distance_between_points)
    pts = []
    for i in range(sections):
        section_pos = (p2 - p1) / (sections + 1) * (i + 1)
        theta_dist = distance_between_points / radius
        for j in range(int(2 * np.pi * radius / distance_between_points)):
            th = theta_dist * j
            pts.append(p1 + section_pos * position_scale + radius * np.cos(th) * ortho + radius * np.sin(th) * normal)
    return np.asarray(pts, dtype=np.float64).reshape(-1, 3)


def bed_bathing_targets(gender: str, hipbone_to_mouth_height: float):
    """`BedBathingEnv.generate_targets` (bed_bathing.py:360-370): (points on the upper arm, points on the forearm), each in
    the COM frame of human link 9 / 11."""
    if gender == "male":
        hmhs = hipbone_to_mouth_height / 0.6
        ul, ur, fl, fr = 0.279, 0.043, 0.257, 0.033
    else:
        hmhs = hipbone_to_mouth_height / 0.54
        ul, ur, fl, fr = 0.264, 0.0355, 0.234, 0.027
    up = capsule_points([0, 0, 0], [0, 0, -ul], ur, 0.03, hmhs)
    fo = capsule_points([0, 0, 0], [0, 0, -fl], fr, 0.03, hmhs)
    return up, fo


def build_bed_bathing(assets_dir: str, robot_type: str = "jaco", gender: str = "male", human_control: bool = False,
                      stage: str = "play", arm_q: Optional[List[float]] = None,
                      base_xy_yaw: Tuple[float, float, float] = (0.0, 0.0, 0.0), verbose: bool = False,
                      new: bool = False, hipbone_to_mouth_height: Optional[float] = None,
                      waist: Tuple[float, float, float] = (0.0, 0.0, 0.0)) -> CompiledScene:
    """BedBathing<Robot>-v0 (bed_bathing.py:155-358).

    `new` = BedBathing<Robot>New-v0 (`__init__.py:122-134`, bed_bathing.py:183-185,256-279): person of height
    `hipbone_to_mouth_height` with revolute waist joints at the drawn `waist` angles, the arm at its preset (:269) plus a per-episode
    draw instead of the settled pose (no 100-step drop in this branch), the whole person static during play.  The arm joints
    stay in the articulation so that the per-episode pose can be written into the record; they are frozen per environment
    through AVG_E_FROZEN (mass 0, world_creation.py:157-161) and carry neither motor nor action.

    stage 'settle': the world as it is during `for _ in range(100): p.stepSimulation()` (bed_bathing.py:286-292) -- human
    base fixed over the bed, right arm (joints 7..13) dynamic under gravity (0, 0, -1), every human joint with a
    velocity motor of 0.1 N m (world_creation.py:164-167), bed parts with friction 5; the robot parked where
    `init_jaco` puts it (world_creation.py:288).
    stage 'play': the world the episode steps in -- the whole human static at the settled pose `arm_q`
    (human_controllable_joint_indices == [] makes every joint static, bed_bathing.py:294-295 + world_creation.py:157-161;
    with `human_control` joints 4..13 stay controllable, i.e. the right arm remains dynamic, starts at `arm_q` and is driven
    by the human half of the action exactly as in ScratchItch: position motors with human_gains 0.05, hard and realistic
    joint limits, env.py:307-337,343-349),
    robot base at the pose chosen by `position_robot_toc` (random_pos x, y and yaw = `base_xy_yaw`, env.py:511-513,
    bed_bathing.py:325), nightstand under it (bed_bathing.py:330-338), gravity off for robot / human / tool (:341-344).
    """
    cfg = CONFIG["bed_bathing"]
    deg = np.deg2rad
    play = stage == "play"
    robot, rs = load_robot(assets_dir, robot_type)
    robot.fixed_base = True
    if play:
        rx, ry, yaw = base_xy_yaw
        robot.base_pos = np.array([-0.85, -0.4, 0.0]) + np.asarray(TOC_POS_OFFSET[("bed_bathing", robot_type)]) \
            + np.array([rx, ry, 0.0])                                                                     # env.py:513, bed_bathing.py:318,325
        robot.base_quat = X.quat_from_euler([0, 0, yaw])
    elif robot_type == "jaco":
        robot.base_pos = np.array([-2.0, -2.0, 0.975]); robot.base_quat = I4.copy()                      # world_creation.py:288
    else:
        robot.base_pos = np.array([-2.0, -2.0, 0.0]); robot.base_quat = I4.copy()                        # world_creation.py:194
    h2m = 0.6 if gender == "male" else 0.54                                                               # bed_bathing.py:196
    if new and hipbone_to_mouth_height is not None:
        h2m = float(hipbone_to_mouth_height)                                                              # bed_bathing.py:184
    assert not (new and (human_control or not play))
    human = create_human(assets_dir, gender, h2m, limit_scale=1.0, static_base=True, new=new)
    human.base_pos = np.array([0.0, 0.0, 0.7]); human.base_quat = X.quat_from_euler([deg(-30), 0, 0])     # bed_bathing.py:203
    human.gravity = np.array([0.0, 0.0, 0.0 if play else -1.0])                                           # :289 / :343
    tool = urdf_to_multibody(os.path.join(assets_dir, "bed_bathing", "wiper.urdf"), REF_TOOL, "wiper")
    tool.fixed_base = False
    bed_friction = 5.0                                                                                    # bed_bathing.py:282-283
    y_offset = -0.53
    m0 = _static_body("mattress0", [ShapeDesc(SHAPE_BOX, np.array([0, 0, 0.15 / 2.0]), I4.copy(), half=np.array([0.88, 1.25, 0.15]) / 2.0,
                                              friction=bed_friction)], [0, y_offset, 0.4], I4)            # :214-216
    m1 = _static_body("mattress1", [ShapeDesc(SHAPE_BOX, np.array([0, 0.7 / 2.0, 0]), I4.copy(), half=np.array([0.88, 0.7, 0.15]) / 2.0,
                                              friction=bed_friction)], [0, 1.25 / 2.0 + y_offset, 0.4 + 0.15 / 2.0],
                      X.quat_from_euler([deg(60), 0, 0]))                                                 # :218-220
    frame = _static_body("bed_frame", _mesh_shapes(os.path.join(assets_dir, "bed", "hospital_bed_frame_vhacd.obj"), [1, 1.2, 1], bed_friction),
                         [0, y_offset + 0.45, 0.42], X.quat_from_euler([np.pi / 2.0, 0, -np.pi / 2.0]))   # :223-227
    plane = _static_body("plane", [ShapeDesc(SHAPE_PLANE, np.zeros(3), I4.copy(), friction=1.0, ref_link=-1)], [0, 0, 0], I4, REF_PLANE)
    mbs = [robot, human, tool, m0, m1, frame]
    if play and robot_type == "jaco":                                                                     # "a nightstand ... for the jaco arm"
        ns = 0.275                                                                                        # :331-338
        nightstand = _static_body("nightstand", _mesh_shapes(os.path.join(assets_dir, "nightstand", "nightstand.obj"), [ns] * 3, 0.5,
                                                             single_hull=True),
                                  np.array([-0.85, 0.12, 0.0]) + np.array([base_xy_yaw[0], base_xy_yaw[1], 0.0]),
                                  X.quat_from_euler([np.pi / 2.0, 0, 0]))
        mbs.append(nightstand)
    mbs.append(plane)

    # -- joint presets and frozen joints ----------------------------------------------------------------------
    q_human = {7: deg(50), 8: deg(-50), 17: deg(-30), 28: deg(-60), 35: deg(-60)}                         # bed_bathing.py:284
    arm = list(range(7, 14))
    if new:                                                                                               # bed_bathing.py:269-271
        q_human = {7: deg(20), 8: deg(-20), 10: deg(-45), 20: deg(-45), 28: deg(-60), 35: deg(-60),
                   0: float(waist[0]), 1: float(waist[1]), 2: float(waist[2])}
    elif play:
        if arm_q is None:
            arm_q = load_settled_arm_q()[gender]
        if arm_q is None:
            raise RuntimeError("no settled arm pose: run tools/settle_bed_bathing.py on a GPU box first")
        for j, v in zip(arm, arm_q):
            q_human[j] = float(v)
    for l in human.links:
        if l.jtype == "revolute":
            q_human[l.ref_index] = float(np.clip(q_human.get(l.ref_index, 0.0), l.lower, l.upper))       # world_creation.py:169
    controllable = list(range(4, 14))                                                                     # bed_bathing.py:294
    frozen_h = {l.ref_index for l in human.links if (play and not human_control) or l.ref_index not in controllable}   # :285 / :294-295
    if new:
        frozen_h = {l.ref_index for l in human.links if l.ref_index not in arm}   # the arm is frozen per environment (AVG_E_FROZEN), not baked
    robot_arm = rs["arm"]
    fingers = rs["fingers"]
    finger_open, tool_pos_offset, tool_euler = TOOL_SETUP[("bed_bathing", robot_type)]                    # bed_bathing.py:320-321,327-328

    def setup_dof(b: DynBody, d: dict) -> None:
        if b.art == 0:
            if b.ref_joint in robot_arm:
                d.update(kp=cfg["robot_gains"], max_force=cfg["robot_forces"], action=robot_arm.index(b.ref_joint))
                d["flags"] |= 2
            elif b.ref_joint in fingers:
                d.update(kp=0.05, max_force=500.0, init_target=finger_open)                               # world_creation.py:328
                d["flags"] |= 2
        elif b.art == 1 and new:
            pass                                                  # static during play: no motor, no action (frozen through AVG_E_FROZEN)
        elif b.art == 1 and play:                                 # human-active ids: take_step's position motors, env.py:337
            slot = controllable.index(b.ref_joint)
            d.update(kp=0.05, max_force=1.0, human_slot=slot, action=7 + slot)
            d["flags"] |= 2 | 4 | 8
        elif b.art == 1:
            d.update(kp=0.0, kd=1.0, max_force=0.1)               # VELOCITY_CONTROL, target 0, force 0.1 (world_creation.py:164-167)
            d["flags"] |= 2

    tool_filtered_robot_links = rs["tool_filtered"]                                                       # world_creation.py:352-354
    bed_filtered_human_links = set(range(28, 42)) | {0, 1, 2, 3}                                          # bed_bathing.py:231-234
    bed_mbs = {3, 4, 5}

    def cross_pair_ok(a: CompiledShape, b: CompiledShape) -> bool:
        ms = {a.mb_index: a, b.mb_index: b}
        if 0 in ms and 2 in ms and ms[0].ref_link in tool_filtered_robot_links:
            return False
        if 1 in ms and (set(ms) & bed_mbs) and ms[1].ref_link in bed_filtered_human_links:
            return False
        return True

    bodies, attach, dofs, n_jdof, n_free, shapes, n_mshape, pairs = _assemble(
        mbs, {0: rs["q_preset"], 1: q_human}, {0: rs["frozen"], 1: frozen_h}, setup_dof, cross_pair_ok, robot, human)
    n_body = len(bodies); n_dof = len(dofs)

    def com_frame(k: int, li: int):
        at = attach[k][li]
        link = mbs[k].link(li)
        p, q = X.tf_mul(at.pos, at.quat, link.inertial_pos, link.inertial_quat)
        return (at.body, p, q)

    tool_pos_offset = np.asarray(tool_pos_offset, float)
    tool_orient_offset = X.quat_from_euler(tool_euler)
    ee = com_frame(0, rs["ee_link"])
    weld_parent = (ee[0],) + X.tf_mul(ee[1], ee[2], tool_pos_offset, tool_orient_offset)
    frames = [
        com_frame(2, 1),            # AVG_F_TOOL_TIP: wiper link 1 (cloth) COM, bed_bathing.py:54,131
        com_frame(2, -1),           # AVG_F_TOOL_BASE
        weld_parent,                # AVG_F_WELD_PARENT
        com_frame(0, rs["torso_link"]),   # AVG_F_TORSO: robot link 0 (PR2: 15) COM, bed_bathing.py:130
        com_frame(1, 3),            # AVG_F_CHEST: human link 3, bed_bathing.py:137
        com_frame(1, 9), com_frame(1, 11), com_frame(1, 13),       # :143-145
    ]
    up, fo = bed_bathing_targets(gender, h2m)
    hp = CONFIG["human_preferences"]
    task_f = np.zeros(32, dtype=np.float32)
    n_target = len(up) + len(fo)
    task_f[:15] = [cfg["distance_weight"], cfg["action_weight"], 0.0, cfg["wiping_reward_weight"],
                   n_target * cfg["task_success_threshold"], hp["velocity_weight"], hp["force_nontarget_weight"],
                   hp["high_forces_weight"], hp["food_hit_weight"], hp["food_velocities_weight"],
                   0.025, 0.0, 10.0, 0.05, 1.0]
    task_f[15] = 4.0                        # getClosestPoints query distance, bed_bathing.py:61
    task_f[16:19] = [-0.4, 0.0, 0.9]        # reference point of the device spatial algebra (float32 conditioning)
    header = dict(task=1, n_body=n_body, n_ebody=0, n_dof=n_dof, n_jdof=n_jdof, n_free=n_free,
                  substeps=5, solver_iters=50,                             # env.py:16, bed_bathing.py:340
                  n_action_robot=7, n_action_human=10 if human_control else 0, n_obs_robot=24,
                  n_obs_human=28 if human_control else 0, human_control=int(human_control),        # bed_bathing.py:19
                  dt=0.02, erp=0.2, lin_damp=0.04, ang_damp=0.04, residual_thr=1e-7, max_vel=100.0,
                  action_scale=0.05, weld_max_force=500.0,
                  weld_body_a=weld_parent[0], weld_body_b=frames[1][0], task_f=task_f)
    scene = CompiledScene(task="bed_bathing", robot_type=robot_type, gender=gender, human_control=human_control,
                          multibodies=mbs, bodies=bodies, attach=attach, shapes=shapes, n_mshape=n_mshape, pairs=pairs,
                          frames=frames, dofs=dofs, header=header, robot_arm_joints=robot_arm,
                          human_joints=arm if (human_control or not play or new) else [], q_human_reset=q_human,
                          tool_offset=(tool_pos_offset, tool_orient_offset))
    scene.mlp_layers = None
    if human_control and play:     # enforce_realistic_human_joint_limits, env.py:343-344,353-387
        from .h5lite import load_keras_dense_stack
        scene.mlp_layers = load_keras_dense_stack(os.path.join(assets_dir, 'realistic_arm_limits_model.h5'))
    scene.targets = (up, fo)
    scene.finger_open = finger_open; scene.robot_spec = rs
    scene.frozen_mask = sum(1 << i for i, b in enumerate(bodies) if b.art == 1) if new else 0
    scene.info = dict(hull_errors=dict(hull_errors), n_pairs=len(pairs), n_shapes=len(shapes), n_mshape=n_mshape, n_target=n_target)
    if verbose:
        print(scene.info)
    return scene
