"""Scene recipes of the Feeding and Drinking tasks (reference `feeding.py:144-331`, `drinking.py:159-335`).

What the reference builds in `reset()` and this module turns into a ModelBlob:

* robot arm + gripper as in the other tasks, but the RIGHT arm drives the tool (`robot_arm='right'`, feeding.py:48) with
  `robot_gains = 0.005` (config.ini:22,30);
* spoon / cup: one rigid link whose collision mesh is a VHACD set of convex hulls (64 / 68 pieces,
  `dinnerware/spoon.urdf`, `cup.urdf`), a free body welded to the gripper (`init_tool`, world_creation.py:331-365)
  -> ONE compound shape with the hulls as children;
* human seated in the wheelchair, everything static but the head chain (joints 24..27), which is dynamic when the id is
  human-active or the episode drew a tremor (feeding.py:244).  The head angles 25..27 are drawn per episode (:243), so the
  chain is always part of the articulation and frozen per environment through AVG_E_FROZEN (mass 0, world_creation.py:157-161);
* Feeding: `table_tall.urdf` (its collision is the table top box) and the bowl (70 hulls), whose position is drawn per episode
  (feeding.py:184) -> an env-static body (pose in AVG_E_EBODY[0]).  Documented deviation: the 0.1 kg bowl is not integrated
  (it rests on the table and only matters when the robot pushes it), DESIGN.md;
* 8 food / 64 water spheres (r = 5 mm, 1 g) under gravity -9.81, everything else gravity-free (feeding.py:284-287);
* `numSubSteps = 2`, 10 solver iterations (feeding.py:289).

Sawyer and Baxter: the reference's task files have no branch for them, so their base placement, start target,
gripper opening and tool offset are DEFINED HERE (see ROBOT_FD) and have no reference counterpart.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np

from . import xform as X
from .human import create_human
from .mbody import DynBody, LinkDesc, MultiBodyDesc, ShapeDesc, SHAPE_BOX, SHAPE_PLANE, SHAPE_SPHERE
from .scene import (CONFIG, CompiledScene, CompiledShape, I4, REF_BOWL, REF_FURNITURE, REF_PLANE, REF_TABLE, REF_TOOL,
                    _assemble, _static_body, expand_compounds, hull_errors, load_robot, urdf_to_multibody)

CONFIG["feeding"] = dict(robot_forces=1.0, robot_gains=0.005, distance_weight=1.0, action_weight=0.01,
                         food_reward_weight=1.0, task_success_threshold=0.75)                       # config.ini:20-26
CONFIG["drinking"] = dict(robot_forces=1.0, robot_gains=0.005, distance_weight=1.0, action_weight=0.01, cup_tilt_weight=0.1,
                          drinking_reward_weight=1.0, task_success_threshold=0.75)                  # config.ini:28-35

PARTICLE_RADIUS = 0.005          # feeding.py:293, drinking.py:294
PARTICLE_MASS = 0.001            # feeding.py:299
BOWL_CENTER = np.array([-0.15, -0.55, 0.75])      # feeding.py:184 (+ U(+-0.05) in x and y per episode)
BOWL_QUAT = X.quat_from_euler([np.pi / 2.0, 0, 0])
TABLE_POS = np.array([0.35, -0.9, 0.0])           # feeding.py:182

# Per (task, robot): gripper open position, tool pos_offset, tool orient_offset (euler), start target of the end effector
# (centre of the +-0.05 box: relative to the bowl for Feeding, absolute for Drinking) and its orientation (euler), IK tolerance,
# TOC pos_offset (None = fixed base).  Jaco / PR2 rows quote the reference; Sawyer / Baxter rows are build-defined (module
# docstring): the tool takes the same WORLD pose as in the Jaco / PR2 recipe (spoon level with its handle towards the robot, cup
# upright), held with the gripper's approach axis (+z of `right_gripper_base`) along the handle / towards the cup's side, and
# the end-effector target is pulled back along y by the extra reach of the parallel gripper.
ROBOT_FD = {
    ("feeding", "jaco"): dict(open=1.33, pos=[0.1, -0.0225, 0.03], euler=[-0.1, -np.pi / 2.0, 0], start=[0, -0.1, 0.4],
                              start_euler=[np.pi / 2.0, 0, np.pi / 2.0], tol=0.01, toc=None),              # feeding.py:276-280
    ("feeding", "pr2"): dict(open=0.03, pos=[0, -0.03, -0.11], euler=[-0.2, 0, 0], start=[0, -0.1, 0.4],
                             start_euler=[np.pi / 2.0, 0, 0], tol=0.03, toc=[0.1, 0.2, 0.0]),               # feeding.py:266-273
    ("feeding", "sawyer"): dict(open=0.01, pos=[0, 0, 0.25], euler=[np.pi - 0.1, 0, 0], start=[0, -0.25, 0.4],
                                start_euler=[-np.pi / 2.0, 0, 0], tol=0.03, toc=[0.0, 0.2, 0.975]),
    ("feeding", "baxter"): dict(open=0.01, pos=[0, 0, 0.25], euler=[np.pi - 0.1, 0, 0], start=[0, -0.25, 0.4],
                                start_euler=[-np.pi / 2.0, 0, 0], tol=0.03, toc=[0.0, 0.2, 0.975]),
    ("drinking", "jaco"): dict(open=0.63, pos=[0.05, -0.005, 0], euler=[0, 0, np.pi / 2.0], start=[-0.2, -0.5, 1.0],
                               start_euler=[0, np.pi / 2.0, 0], tol=0.01, toc=None),                        # drinking.py:272-276
    ("drinking", "pr2"): dict(open=0.45, pos=[-0.01, 0, -0.05], euler=[np.pi / 2.0, 0, 0], start=[-0.2, -0.5, 1.0],
                              start_euler=[0, 0, 0], tol=0.03, toc=[0.2, 0.2, 0.0]),                        # drinking.py:263-270
    ("drinking", "sawyer"): dict(open=0.02, pos=[0, 0.06, 0.19], euler=[np.pi, 0, 0], start=[-0.2, -0.64, 1.0],
                                 start_euler=[-np.pi / 2.0, 0, 0], tol=0.03, toc=[0.0, 0.2, 0.975]),
    ("drinking", "baxter"): dict(open=0.02, pos=[0, 0.06, 0.19], euler=[np.pi, 0, 0], start=[-0.2, -0.64, 1.0],
                                 start_euler=[-np.pi / 2.0, 0, 0], tol=0.03, toc=[0.0, 0.2, 0.975]),
}


def start_target(task: str, robot_type: str) -> Tuple[np.ndarray, np.ndarray]:
    """Centre of the start-target box of the end effector and its orientation (feeding.py:276-277, drinking.py:272-273)."""
    r = ROBOT_FD[(task, robot_type)]
    c = np.asarray(r["start"], float) + (BOWL_CENTER if task == "feeding" else 0.0)
    return c, X.quat_from_euler(r["start_euler"])


def particle_grid(task: str) -> np.ndarray:
    """Offsets of the freshly created particles from the tool's base position, world axes (feeding.py:301-305,
    drinking.py:302-306), in the reference's creation order (= the order of `self.foods` / `self.waters`)."""
    r = PARTICLE_RADIUS
    if task == "feeding":
        return np.array([[i * 2 * r - 0.005, j * 2 * r, k * 2 * r + 0.02] for i in range(2) for j in range(2) for k in range(2)])
    return np.array([[i * 2 * r - 0.02, j * 2 * r - 0.02, k * 2 * r + 0.075] for i in range(4) for j in range(4) for k in range(4)])


def build_feeding_drinking(assets_dir: str, task: str = "feeding", robot_type: str = "jaco", gender: str = "male",
                           human_control: bool = False, base_xy_yaw: Tuple[float, float, float] = (0.0, 0.0, 0.0),
                           verbose: bool = False, new: bool = False, hipbone_to_mouth_height: float | None = None,
                           waist: Tuple[float, float, float] = (0.0, 0.0, 0.0)) -> CompiledScene:
    """Feeding<Robot>[Human]-v0 / Drinking<Robot>[Human]-v0.  `base_xy_yaw` = random_pos x, y and yaw of the robot base chosen by
    `position_robot_toc` (env.py:511-513; PR2, and the build-defined Sawyer / Baxter placements).
    `new` = <Task><Robot>New-v0 (`__init__.py:206-218,290-302`): person of height `hipbone_to_mouth_height` (feeding.py:171) with
    revolute waist joints held at the drawn `waist` angles (feeding.py:233); the whole person is static in these ids (:235)."""
    assert task in ("feeding", "drinking")
    cfg = CONFIG[task]
    rec = ROBOT_FD[(task, robot_type)]
    deg = np.deg2rad
    robot, rs = load_robot(assets_dir, robot_type, arm="right")
    robot.fixed_base = True
    if rec["toc"] is None:
        robot.base_pos = np.array([-0.35, -0.3, 0.36])                                            # feeding.py:188, drinking.py:195
        robot.base_quat = np.array([0.0, 0.0, -0.7071067811865475, 0.7071067811865476])
    else:
        robot.base_pos = np.array([-0.85, -0.4, 0.0]) + np.asarray(rec["toc"], float) + np.array([base_xy_yaw[0], base_xy_yaw[1], 0.0])   # env.py:513
        robot.base_quat = X.quat_from_euler([0, 0, base_xy_yaw[2]])
    h2m = 0.6 if gender == "male" else 0.54                                                       # feeding.py:174
    if new and hipbone_to_mouth_height is not None:
        h2m = float(hipbone_to_mouth_height)                                                      # feeding.py:171
    human = create_human(assets_dir, gender, h2m, limit_scale=1.0, static_base=True, new=new)
    human.base_pos = np.array([0, 0.03, 0.89 - 0.23725 if gender == "male" else 0.86 - 0.225])    # feeding.py:245
    human.compound_links = {27}                                                                   # head: 8 / 9 VHACD hulls
    tool = urdf_to_multibody(os.path.join(assets_dir, "dinnerware", "spoon.urdf" if task == "feeding" else "cup.urdf"), REF_TOOL,
                             "spoon" if task == "feeding" else "cup")                             # world_creation.py:341-343
    tool.fixed_base = False
    tool.compound_links = {-1}
    chair = urdf_to_multibody(os.path.join(assets_dir, "wheelchair", "wheelchair.urdf"), REF_FURNITURE, "wheelchair")
    chair.base_pos = np.array([0.0, 0.09, -0.01])                                                 # world_creation.py:49
    chair.base_quat = X.quat_from_euler([np.pi / 2.0, 0, -np.pi / 2.0 - 0.05])
    plane = _static_body("plane", [ShapeDesc(SHAPE_PLANE, np.zeros(3), I4.copy(), friction=1.0, ref_link=-1)], [0, 0, 0], I4, REF_PLANE)
    mbs = [robot, human, tool, chair]
    n_ebody = 0
    if task == "feeding":
        table = urdf_to_multibody(os.path.join(assets_dir, "table", "table_tall.urdf"), REF_TABLE, "table")   # feeding.py:182
        table.base_pos = TABLE_POS.copy(); table.fixed_base = True
        bowl = urdf_to_multibody(os.path.join(assets_dir, "dinnerware", "bowl.urdf"), REF_BOWL, "bowl")       # feeding.py:185
        bowl.fixed_base = True; bowl.env_static = 0; bowl.compound_links = {-1}
        bowl.base_pos = np.zeros(3); bowl.base_quat = I4.copy()
        mbs += [table, bowl]
        n_ebody = 1
    mbs.append(plane)
    i_tool = 2

    # -- joint presets and frozen joints (feeding.py:242-244) ---------------------------------------------------
    q_human = {10: deg(-90), 20: deg(-90), 28: deg(-90), 31: deg(80), 35: deg(-90), 38: deg(80)}
    if new:
        q_human.update({0: float(waist[0]), 1: float(waist[1]), 2: float(waist[2])})              # feeding.py:233
    for l in human.links:
        if l.jtype == "revolute":
            q_human[l.ref_index] = float(np.clip(q_human.get(l.ref_index, 0.0), l.lower, l.upper))      # world_creation.py:172
    controllable = [24, 25, 26, 27]                                                                # feeding.py:219
    frozen_h = {l.ref_index for l in human.links if l.ref_index not in controllable}
    robot_arm = rs["arm"]
    fingers = rs["fingers"]
    signs = rs.get("finger_signs", [1.0] * len(fingers))                                           # world_creation.py:313-320
    finger_open = float(rec["open"])

    def setup_dof(b: DynBody, d: dict) -> None:
        if b.art == 0:
            if b.ref_joint in robot_arm:
                d.update(kp=cfg["robot_gains"], max_force=cfg["robot_forces"], action=robot_arm.index(b.ref_joint))
                d["flags"] |= 2
            elif b.ref_joint in fingers:
                d.update(kp=0.05, max_force=500.0, init_target=finger_open * signs[fingers.index(b.ref_joint)])   # world_creation.py:328
                d["flags"] |= 2
        elif b.art == 1:                                           # head chain: take_step's position motors (env.py:337) with
            slot = controllable.index(b.ref_joint)                 # human_gains = 0.005 (feeding.py:48), force 1 * strength
            d.update(kp=0.005, max_force=1.0, human_slot=slot)
            d["flags"] |= 2 | 4 | 8
            if human_control:
                d["action"] = 7 + slot

    tool_filtered_robot_links = rs["tool_filtered"]                                                # world_creation.py:359-361

    def cross_pair_ok(a: CompiledShape, b: CompiledShape) -> bool:
        ms = {a.mb_index: a, b.mb_index: b}
        return not (0 in ms and i_tool in ms and ms[0].ref_link in tool_filtered_robot_links)

    bodies, attach, dofs, n_jdof, n_free, shapes, n_mshape, pairs = _assemble(
        mbs, {0: rs["q_preset"], 1: q_human}, {0: rs["frozen"], 1: frozen_h}, setup_dof, cross_pair_ok, robot, human)
    n_body = len(bodies); n_dof = len(dofs)
    cshapes, opairs = expand_compounds(shapes, pairs)

    def com_frame(k: int, li: int):
        at = attach[k][li]
        link = mbs[k].link(li)
        p, q = X.tf_mul(at.pos, at.quat, link.inertial_pos, link.inertial_quat)
        return (at.body, p, q)

    tool_pos_offset = np.asarray(rec["pos"], float)
    tool_orient_offset = X.quat_from_euler(rec["euler"])
    ee = com_frame(0, rs["ee_link"])
    weld_parent = (ee[0],) + X.tf_mul(ee[1], ee[2], tool_pos_offset, tool_orient_offset)
    tool_base = com_frame(i_tool, -1)
    chest = com_frame(1, 3)
    frames = [
        tool_base,                  # AVG_F_TOOL_TIP: the tool's base (getBasePositionAndOrientation(spoon), feeding.py:66,125)
        tool_base,                  # AVG_F_TOOL_BASE
        weld_parent,                # AVG_F_WELD_PARENT
        com_frame(0, rs["torso_link"]),   # AVG_F_TORSO: robot link 0 (PR2: 15), feeding.py:124
        chest,                      # AVG_F_CHEST: human link 3, feeding.py:130
        chest, chest, chest,        # AVG_F_SHOULDER / ELBOW / WRIST: not used by these tasks
        com_frame(1, 27),           # AVG_F_HEAD: human link 27, feeding.py:134,346
    ]
    hp = CONFIG["human_preferences"]
    n_particle = 8 if task == "feeding" else 64
    task_f = np.zeros(32, dtype=np.float32)
    task_f[:15] = [cfg["distance_weight"], cfg["action_weight"], 0.0, 0.0, n_particle * cfg["task_success_threshold"],
                   hp["velocity_weight"], hp["force_nontarget_weight"], hp["high_forces_weight"], hp["food_hit_weight"],
                   hp["food_velocities_weight"], 0.0, 0.0, 10.0, 0.005, 1.0]                      # human_gains 0.005: feeding.py:48
    task_f[16:19] = [-0.3, -0.3, 0.9]       # reference point of the device spatial algebra (float32 conditioning)
    task_f[19:22] = [0.0, -0.11 if gender == "male" else -0.1, 0.03]                              # mouth_pos, feeding.py:253
    if task == "feeding":
        task_f[22:27] = [0.02, 20.0, -5.0, 0.5, cfg["food_reward_weight"]]                        # feeding.py:102-113
    else:
        task_f[22:27] = [0.03, 10.0, -1.0, 0.5, cfg["drinking_reward_weight"]]                    # drinking.py:114-126
        task_f[27] = cfg["cup_tilt_weight"]
        task_f[28] = 1.0 if robot_type == "jaco" else -1.0                                        # drinking.py:72
        task_f[29:32] = [0.05, -0.055, 0.07]                                                      # drinking.py:112,278-279
    head_mask = 0
    for i, b in enumerate(bodies):
        if b.art == 1:
            head_mask |= 1 << i
    header = dict(task=2 if task == "feeding" else 3, n_body=n_body, n_ebody=n_ebody, n_dof=n_dof, n_jdof=n_jdof, n_free=n_free,
                  substeps=5, solver_iters=10, n_internal=2,                                      # env.py:16, feeding.py:289
                  n_action_robot=7, n_action_human=4 if human_control else 0,
                  n_obs_robot=25, n_obs_human=23 if human_control else 0, human_control=int(human_control),   # feeding.py:18
                  dt=0.01, erp=0.2, lin_damp=0.04, ang_damp=0.04, residual_thr=1e-7, max_vel=100.0,
                  action_scale=0.05, weld_max_force=500.0,
                  weld_body_a=weld_parent[0], weld_body_b=tool_base[0], task_f=task_f,
                  n_particle=n_particle, p_mass=PARTICLE_MASS, p_gravity=[0.0, 0.0, -9.81],
                  tool_body=tool_base[0], head_frozen_mask=head_mask)
    scene = CompiledScene(task=task, robot_type=robot_type, gender=gender, human_control=human_control,
                          multibodies=mbs, bodies=bodies, attach=attach, shapes=shapes, n_mshape=n_mshape, pairs=pairs,
                          frames=frames, dofs=dofs, header=header, robot_arm_joints=robot_arm,
                          human_joints=controllable, q_human_reset=q_human,
                          tool_offset=(tool_pos_offset, tool_orient_offset), cshapes=cshapes, opairs=opairs)
    scene.finger_open = finger_open; scene.finger_signs = signs; scene.robot_spec = rs
    scene.mlp_layers = None
    # the particle template: a sphere whose contact-breaking threshold is 0.02 x its radius (btCollisionShape::
    # getContactBreakingThreshold with the angular-motion disc of a centred sphere) [UPSTREAM-BULLET]
    scene.particle = ShapeDesc(SHAPE_SPHERE, np.zeros(3), I4.copy(), radius=PARTICLE_RADIUS, friction=0.5, ref_link=-1)
    scene.particle_thr = 0.02 * PARTICLE_RADIUS
    scene.info = dict(hull_errors=dict(hull_errors), n_pairs=len(pairs), n_opairs=len(opairs), n_shapes=len(shapes),
                      n_mshape=n_mshape, n_cshape=len(cshapes), n_body=n_body, n_dof=n_dof)
    if verbose:
        print(scene.info)
    return scene
