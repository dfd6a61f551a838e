"""Procedural capsule human, restating reference `human_creation.py:57-301` (`HumanCreation.create_human`).

Produces a `MultiBodyDesc` whose link indices follow PyBullet's depth-first renumbering of the creation order, i.e.
the legend at `human_creation.py:5-45` (right arm = joints 7-13, left arm 17-23, neck 24, head 25-27, legs 28-41).
Only dimensions, offsets, masses, axes, limits and the self-collision rule are taken from the reference; the data
structure is ours.
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import numpy as np

from . import xform as X
from .mbody import (LinkDesc, MultiBodyDesc, ShapeDesc, SHAPE_CAPSULE, SHAPE_SPHERE, SHAPE_HULL)
from .meshes import load_obj_groups, prepare_hull

I4 = np.array([0.0, 0, 0, 1])
CONFIG_HUMAN = {  # reference config.ini:46-54
    "male": dict(mass=78.4, radius_scale=1.0, height_scale=1.0),
    "female": dict(mass=62.5, radius_scale=1.0, height_scale=1.0),
}


def _cap(radius, length, pos=(0, 0, 0), quat=I4):
    return ShapeDesc(SHAPE_CAPSULE, np.asarray(pos, float), np.asarray(quat, float), radius=radius,
                     half=np.array([0.0, 0.0, length / 2.0]))


def _sph(radius, pos=(0, 0, 0)):
    return ShapeDesc(SHAPE_SPHERE, np.asarray(pos, float), I4.copy(), radius=radius)


def human_dims(gender: str, hipbone_to_mouth_height: float) -> dict:
    """Shape sizes and link offsets, `human_creation.py:70-115` (male) / `:116-161` (female)."""
    c = CONFIG_HUMAN[gender]
    rs = c["radius_scale"]
    qx = X.quat_from_euler([0, np.pi / 2.0, 0])
    qf = X.quat_from_euler([np.pi / 2.0, 0, 0])
    if gender == "male":
        hs = c["height_scale"] * hipbone_to_mouth_height / 0.6
        d = dict(
            chest=_cap(0.127 * rs, 0.056, quat=qx),
            right_shoulders=_cap(0.106 * rs, 0.253 / 8, pos=[-0.253 / 2.5 + 0.253 / 16, 0, 0], quat=qx),
            left_shoulders=_cap(0.106 * rs, 0.253 / 8, pos=[0.253 / 2.5 - 0.253 / 16, 0, 0], quat=qx),
            neck=_cap(0.06 * rs, 0.124 * hs, pos=[0, 0, (0.2565 - 0.1415 - 0.025) * hs]),
            upperarm=_cap(0.043 * rs, 0.279 * hs, pos=[0, 0, -0.279 / 2.0 * hs]),
            forearm=_cap(0.033 * rs, 0.257 * hs, pos=[0, 0, -0.257 / 2.0 * hs]),
            hand=_sph(0.043 * rs, pos=[0, 0, -0.043 * rs]),
            waist=_cap(0.1205 * rs, 0.049, quat=qx),
            hips=_cap(0.1335 * rs, 0.094, pos=[0, 0, -0.08125 * hs], quat=qx),
            thigh=_cap(0.08 * rs, 0.424 * hs, pos=[0, 0, -0.424 / 2.0 * hs]),
            shin=_cap(0.05 * rs, 0.403 * hs, pos=[0, 0, -0.403 / 2.0 * hs]),
            foot=_cap(0.05 * rs, 0.215 * hs, pos=[0, -0.1, -0.025 * rs], quat=qf),
            head_file="BaseHeadMeshes_v5_male_cropped_reduced_compressed_vhacd.obj",
            head_pos=[0.09, 0.08, -0.07 + 0.01],
            chest_p=[0, 0, 0.156 * hs], shoulders_p=[0, 0, 0.1415 / 2 * hs], neck_p=[0, 0, 0.1515 * hs],
            head_p=[0, 0, (0.399 - 0.1415 - 0.1205) * hs],
            right_upperarm_p=[-0.106 * rs - 0.073, 0, 0], left_upperarm_p=[0.106 * rs + 0.073, 0, 0],
            forearm_p=[0, 0, -0.279 * hs], hand_p=[0, 0, -(0.033 * rs + 0.257 * hs)],
            waist_p=[0, 0, 0.08125 * hs],
            right_thigh_p=[-0.08 * rs - 0.009, 0, -0.08125 * hs], left_thigh_p=[0.08 * rs + 0.009, 0, -0.08125 * hs],
            shin_p=[0, 0, -0.424 * hs], foot_p=[0, 0, -0.403 * hs - 0.025],
            limb_dims={9: (0.279, 0.043), 11: (0.257, 0.033)},   # scratch_itch.py:277-278
        )
    else:
        hs = c["height_scale"] * hipbone_to_mouth_height / 0.54
        d = dict(
            chest=_cap(0.127 * rs, 0.01, quat=qx),
            right_shoulders=_cap(0.092 * rs, 0.225 / 8, pos=[-0.225 / 2.5 + 0.225 / 16, 0, 0], quat=qx),
            left_shoulders=_cap(0.092 * rs, 0.225 / 8, pos=[0.225 / 2.5 - 0.225 / 16, 0, 0], quat=qx),
            neck=_cap(0.05 * rs, 0.121 * hs, pos=[0, 0, (0.2565 - 0.1415 - 0.025) * hs]),
            upperarm=_cap(0.0355 * rs, 0.264 * hs, pos=[0, 0, -0.264 / 2.0 * hs]),
            forearm=_cap(0.027 * rs, 0.234 * hs, pos=[0, 0, -0.234 / 2.0 * hs]),
            hand=_sph(0.0355 * rs, pos=[0, 0, -0.0355 * rs]),
            waist=_cap(0.11 * rs, 0.009, quat=qx),
            hips=_cap(0.127 * rs, 0.117, pos=[0, 0, -0.15 / 2 * hs], quat=qx),
            thigh=_cap(0.0775 * rs, 0.391 * hs, pos=[0, 0, -0.391 / 2.0 * hs]),
            shin=_cap(0.045 * rs, 0.367 * hs, pos=[0, 0, -0.367 / 2.0 * hs]),
            foot=_cap(0.045 * rs, 0.195 * hs, pos=[0, -0.09, -0.0225 * rs], quat=qf),
            head_file="BaseHeadMeshes_v5_female_cropped_reduced_compressed_vhacd.obj",
            head_pos=[-0.089, -0.09, -0.07],
            chest_p=[0, 0, 0.15 * hs], shoulders_p=[0, 0, 0.132 / 2 * hs], neck_p=[0, 0, 0.132 * hs],
            head_p=[0, 0, 0.12 * hs],
            right_upperarm_p=[-0.092 * rs - 0.067, 0, 0], left_upperarm_p=[0.092 * rs + 0.067, 0, 0],
            forearm_p=[0, 0, -0.264 * hs], hand_p=[0, 0, -(0.027 * rs + 0.234 * hs)],
            waist_p=[0, 0, 0.15 / 2 * hs],
            right_thigh_p=[-0.0775 * rs - 0.0145, 0, -0.15 / 2 * hs], left_thigh_p=[0.0775 * rs + 0.0145, 0, -0.15 / 2 * hs],
            shin_p=[0, 0, -0.391 * hs], foot_p=[0, 0, -0.367 * hs - 0.045 / 2],
            limb_dims={9: (0.264, 0.0355), 11: (0.234, 0.027)},  # scratch_itch.py:279-280
        )
    d["mass"] = c["mass"]
    return d


def create_human(assets_dir: str, gender: str, hipbone_to_mouth_height: float, limit_scale: float = 1.0,
                 static_base: bool = True, new: bool = False) -> MultiBodyDesc:
    d = human_dims(gender, hipbone_to_mouth_height)
    m = d["mass"]
    deg = np.deg2rad
    J0 = [0.0, 0.0, 0.0]

    # creation-order table: (mass fraction, shape(s), position, parent (1-based, 0=base), type, axis, lower, upper)
    rows: List[tuple] = []

    def add(frac, shape, pos, parent, jt, axis, lo=0.0, hi=0.0):
        rows.append((frac, shape, pos, parent, jt, axis, lo, hi))

    # waist and chest, human_creation.py:176-194
    wt = "revolute" if new else "fixed"
    wl = [deg(-30), deg(-30), deg(-30)] if new else [0, 0, 0]
    wu = [deg(75), deg(30), deg(30)] if new else [0, 0, 0]
    wa = [[1, 0, 0], [0, 1, 0], [0, 0, 1]] if new else [J0, J0, J0]
    add(0, None, d["waist_p"], 0, wt, wa[0], wl[0], wu[0])
    add(0, None, J0, 1, wt, wa[1], wl[1], wu[1])
    add(0.13, d["waist"], J0, 2, wt, wa[2], wl[2], wu[2])
    add(0.1, d["chest"], d["chest_p"], 3, "fixed", J0)
    # shoulders, neck, head, human_creation.py:197-209
    head_shapes = "HEAD"
    ls = limit_scale
    add(0, None, d["shoulders_p"], 4, "fixed", J0)
    add(0, None, d["shoulders_p"], 5, "fixed", J0)
    add(0.05, d["right_shoulders"], J0, 6, "fixed", J0)
    add(0, None, d["shoulders_p"], 4, "fixed", J0)
    add(0, None, d["shoulders_p"], 8, "fixed", J0)
    add(0.05, d["left_shoulders"], J0, 9, "fixed", J0)
    add(0.01, d["neck"], d["neck_p"], 4, "revolute", [1, 0, 0], deg(-10) * ls, deg(20) * ls)
    add(0, None, d["head_p"], 11, "revolute", [1, 0, 0], deg(-50) * ls, deg(50) * ls)
    add(0, None, J0, 12, "revolute", [0, 1, 0], deg(-34) * ls, deg(34) * ls)
    add(0.07, head_shapes, J0, 13, "revolute", [0, 0, 1], deg(-70) * ls, deg(70) * ls)
    # right arm, human_creation.py:212-228
    arm_frac = [0, 0, 0.033, 0, 0.019, 0, 0.0065]
    arm_shapes = [None, None, d["upperarm"], None, d["forearm"], None, d["hand"]]
    arm_axes = [[0, 1, 0], [1, 0, 0], [0, 0, 1], [1, 0, 0], [0, 0, 1], [1, 0, 0], [0, 1, 0]]
    r_lo = [5, -188, -90, -128, -90, -81, -27]
    r_hi = [198, 61, 90, 0, 90, 90, 47]
    r_pos = [d["right_upperarm_p"], J0, J0, d["forearm_p"], J0, d["hand_p"], J0]
    r_par = [7, 15, 16, 17, 18, 19, 20]
    for k in range(7):
        add(arm_frac[k], arm_shapes[k], r_pos[k], r_par[k], "revolute", arm_axes[k], deg(r_lo[k]) * ls, deg(r_hi[k]) * ls)
    # left arm, human_creation.py:230-246
    l_lo = [-198, -188, -90, -128, -90, -81, -47]
    l_hi = [-5, 61, 90, 0, 90, 90, 27]
    l_pos = [d["left_upperarm_p"], J0, J0, d["forearm_p"], J0, d["hand_p"], J0]
    l_par = [10, 22, 23, 24, 25, 26, 27]
    for k in range(7):
        add(arm_frac[k], arm_shapes[k], l_pos[k], l_par[k], "revolute", arm_axes[k], deg(l_lo[k]) * ls, deg(l_hi[k]) * ls)
    # legs, human_creation.py:248-274 (leg limits are not scaled by limit_scale)
    leg_frac = [0, 0, 0.105, 0.0475, 0, 0, 0.014]
    leg_shapes = [None, None, d["thigh"], d["shin"], None, None, d["foot"]]
    leg_axes = [[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]]
    rl_lo = [-127, -40, -45, 0, -35, -23, -43]; rl_hi = [30, 45, 40, 130, 38, 24, 35]
    ll_lo = [-127, -45, -40, 0, -35, -24, -35]; ll_hi = [30, 40, 45, 130, 38, 23, 43]
    rl_pos = [d["right_thigh_p"], J0, J0, d["shin_p"], d["foot_p"], J0, J0]
    ll_pos = [d["left_thigh_p"], J0, J0, d["shin_p"], d["foot_p"], J0, J0]
    rl_par = [0, 29, 30, 31, 32, 33, 34]
    ll_par = [0, 36, 37, 38, 39, 40, 41]
    for k in range(7):
        add(leg_frac[k], leg_shapes[k], rl_pos[k], rl_par[k], "revolute", leg_axes[k], deg(rl_lo[k]), deg(rl_hi[k]))
    for k in range(7):
        add(leg_frac[k], leg_shapes[k], ll_pos[k], ll_par[k], "revolute", leg_axes[k], deg(ll_lo[k]), deg(ll_hi[k]))

    # depth-first renumbering, children in creation order (what PyBullet's URDF pipeline does)
    n = len(rows)
    children: Dict[int, List[int]] = {k: [] for k in range(-1, n)}
    for i, r in enumerate(rows):
        children[r[3] - 1].append(i)
    order: List[int] = []

    def dfs(k):
        for ch in children[k]:
            order.append(ch)
            dfs(ch)

    dfs(-1)
    new_index = {old: new_i for new_i, old in enumerate(order)}
    new_index[-1] = -1

    head_hulls = None
    links: List[LinkDesc] = []
    for old in order:
        frac, shape, pos, parent, jt, axis, lo, hi = rows[old]
        idx = new_index[old]
        shapes: List[ShapeDesc] = []
        if isinstance(shape, str) and shape == "HEAD":
            if head_hulls is None:
                groups = load_obj_groups(os.path.join(assets_dir, "head_female_male", d["head_file"]))
                head_hulls = []
                hq = X.quat_from_euler([np.pi / 2.0, 0, 0])
                for g in groups:
                    v, pl, _ = prepare_hull(g * 0.89)          # meshScale, human_creation.py:97
                    head_hulls.append(ShapeDesc(SHAPE_HULL, np.asarray(d["head_pos"], float), hq.copy(), verts=v, planes=pl))
            shapes = head_hulls
        elif shape is not None:
            shapes = [ShapeDesc(shape.kind, shape.pos.copy(), shape.quat.copy(), radius=shape.radius, half=shape.half.copy())]
        for s in shapes:
            s.ref_link = idx
        ax = np.asarray(axis, float)
        links.append(LinkDesc(ref_index=idx, parent=new_index[parent - 1], jtype=jt, pos=np.asarray(pos, float),
                              quat=I4.copy(), axis=ax, mass=m * frac, inertial_pos=np.zeros(3), inertial_quat=I4.copy(),
                              shapes=shapes, lower=float(lo), upper=float(hi),
                              limit_enforced=(jt == "revolute" and lo <= hi)))
    hips = d["hips"]
    base_shape = ShapeDesc(hips.kind, hips.pos.copy(), hips.quat.copy(), radius=hips.radius, half=hips.half.copy(), ref_link=-1)
    base = LinkDesc(ref_index=-1, parent=-1, jtype="base", pos=np.zeros(3), quat=I4.copy(), axis=np.zeros(3),
                    mass=0.0 if static_base else m * 0.14, inertial_pos=np.zeros(3), inertial_quat=I4.copy(),
                    shapes=[base_shape])
    mb = MultiBodyDesc(name=f"human_{gender}", ref_body=1, base=base, links=links, fixed_base=static_base)
    mb.dims = d
    return mb


def human_self_collision_enabled(i: int, j: int, num_joints: int = 42) -> bool:
    """Self-collision filter matrix, `human_creation.py:279-294` (last write wins, as with setCollisionFilterPair)."""
    state = False

    def rule(a_range, b_list):
        nonlocal state
        if i in a_range and j in b_list:
            state = True
        if j in a_range and i in b_list:
            state = True

    rule(range(7, 14), list(range(-1, 4)) + list(range(14, num_joints)))
    rule(range(17, 24), list(range(-1, 14)) + list(range(24, num_joints)))
    rule(range(28, 35), [-1] + list(range(4, 28)) + list(range(35, num_joints)))
    rule(range(35, num_joints), [-1] + list(range(4, 35)))
    return state
