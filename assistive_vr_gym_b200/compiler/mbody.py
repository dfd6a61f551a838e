"""Multibody description used by the model compiler, and its reduction to dynamic bodies.

`MultiBodyDesc` holds a body the way PyBullet sees it after `loadURDF` / `createMultiBody` (reference
`world_creation.py:274-293`, `human_creation.py:275`): one link per joint in PyBullet index order, inertial (COM)
frames, masses, collision shapes. `reduce_bodies` then produces what the device kernels integrate:

* links behind `fixed` joints are merged into their parent (composite mass / COM / inertia); PyBullet keeps them as
  0-DoF links of the same btMultiBody, which is dynamically identical;
* links frozen by the reference (`changeDynamics(mass=0)` on every non-controllable human joint, reference
  `world_creation.py:157-161`; Bullet's ABA then yields zero joint acceleration for them) are
  merged into the static world at their reset pose;
* every remaining body gets a frame at its composite COM aligned with the principal axes, so inertia is diagonal.

Inertia follows Bullet's rule for URDF bodies loaded without `URDF_USE_INERTIA_FROM_FILE`: the box inertia of the
link's collision AABB in its inertial frame, `I = m/12 (ly²+lz², lx²+lz², lx²+ly²)`
(`btCompoundShape::calculateLocalInertia`) [UPSTREAM-BULLET, unverified].
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import xform as X

SHAPE_SPHERE, SHAPE_CAPSULE, SHAPE_BOX, SHAPE_CYLINDER, SHAPE_HULL, SHAPE_PLANE, SHAPE_COMPOUND = 0, 1, 2, 3, 4, 5, 6
JOINT_REVOLUTE, JOINT_PRISMATIC, JOINT_FREE = 0, 1, 2


@dataclass
class ShapeDesc:
    kind: int
    pos: np.ndarray                 # in the owning LINK frame (URDF link frame)
    quat: np.ndarray
    radius: float = 0.0             # sphere/capsule/cylinder radius
    half: np.ndarray = field(default_factory=lambda: np.zeros(3))   # box half extents; capsule/cylinder: half[2] = half length
    verts: Optional[np.ndarray] = None      # hull vertices (shape frame)
    planes: Optional[np.ndarray] = None     # hull planes (n, d)
    friction: float = 0.5
    ref_link: int = -1              # PyBullet link index this shape reports as
    children: Optional[list] = None # SHAPE_COMPOUND: the convex pieces (ShapeDesc, poses in the same frame as this shape's pos/quat = identity)

    def local_aabb(self, margin_in_aabb: bool = True) -> Tuple[np.ndarray, np.ndarray]:
        """AABB (min, max) in the shape's own frame."""
        if self.kind == SHAPE_SPHERE:
            h = np.full(3, self.radius)
        elif self.kind == SHAPE_CAPSULE:
            h = np.array([self.radius, self.radius, self.half[2] + self.radius])
        elif self.kind == SHAPE_BOX:
            h = self.half.copy()
        elif self.kind == SHAPE_CYLINDER:
            h = np.array([self.radius, self.radius, self.half[2]])
        elif self.kind == SHAPE_HULL:
            m = 0.001 if margin_in_aabb else 0.0
            return self.verts.min(0) - m, self.verts.max(0) + m
        elif self.kind == SHAPE_COMPOUND:
            lo = np.full(3, np.inf); hi = np.full(3, -np.inf)
            for c in self.children:
                R = X.quat_to_mat(c.quat)
                clo, chi = c.local_aabb(margin_in_aabb)
                cc = R @ (0.5 * (clo + chi)) + c.pos; ch = np.abs(R) @ (0.5 * (chi - clo))
                lo = np.minimum(lo, cc - ch); hi = np.maximum(hi, cc + ch)
            return lo, hi
        else:                       # ground plane = the reference's 30 x 30 x 10 box centred at z = -5 (plane.urdf)
            return np.array([-15.0, -15.0, -10.0]), np.array([15.0, 15.0, 0.0])
        return -h, h


@dataclass
class LinkDesc:
    ref_index: int                  # PyBullet link index (-1 = base)
    parent: int                     # PyBullet index of the parent link (-1 = base); ignored for the base
    jtype: str                      # 'fixed' | 'revolute' | 'prismatic' | 'base'
    pos: np.ndarray                 # parent LINK frame -> this link frame at q = 0
    quat: np.ndarray
    axis: np.ndarray                # in this link's frame
    mass: float
    inertial_pos: np.ndarray
    inertial_quat: np.ndarray
    shapes: List[ShapeDesc] = field(default_factory=list)
    lower: float = 0.0              # as reported by getJointInfo
    upper: float = -1.0
    limit_enforced: bool = False    # Bullet adds a joint-limit constraint (revolute/prismatic with lower <= upper)
    inertia_diag: Optional[np.ndarray] = None   # None -> from collision AABB (Bullet default)
    name: str = ""
    damping: float = 0.0            # URDF <dynamics damping>: btMultibodyLink::m_jointDamping


@dataclass
class MultiBodyDesc:
    name: str
    ref_body: int                   # id used in contact reports (0 robot, 1 human, 2 tool, 3 furniture, 4 plane)
    base: LinkDesc
    links: List[LinkDesc]           # links[i].ref_index == i
    base_pos: np.ndarray = field(default_factory=lambda: np.zeros(3))   # world pose of the base LINK frame
    base_quat: np.ndarray = field(default_factory=lambda: np.array([0.0, 0, 0, 1]))
    fixed_base: bool = True
    gravity: np.ndarray = field(default_factory=lambda: np.zeros(3))

    def link(self, i: int) -> LinkDesc:
        return self.base if i < 0 else self.links[i]

    # ---- plain numpy kinematics (used by reset/IK and by the tests) ---------------------------------------
    def link_frames(self, q: Dict[int, float]) -> Dict[int, Tuple[np.ndarray, np.ndarray]]:
        """World pose of every LINK frame given joint positions {pybullet joint index: value}."""
        out = {-1: (np.asarray(self.base_pos, float), np.asarray(self.base_quat, float))}
        for l in self.links:
            pp, pq = out[l.parent]
            p, r = X.tf_mul(pp, pq, l.pos, l.quat)
            v = q.get(l.ref_index, 0.0)
            if l.jtype == "revolute":
                r = X.quat_normalize(X.quat_mul(r, X.quat_from_axis_angle(l.axis, v)))
            elif l.jtype == "prismatic":
                p = p + X.quat_rotate(r, l.axis * v)
            out[l.ref_index] = (p, r)
        return out

    def com_frames(self, q: Dict[int, float]) -> Dict[int, Tuple[np.ndarray, np.ndarray]]:
        """World pose of every inertial (COM) frame — what `getLinkState(...)[0:2]` returns."""
        lf = self.link_frames(q)
        out = {}
        for i, (p, r) in lf.items():
            l = self.link(i)
            out[i] = X.tf_mul(p, r, l.inertial_pos, l.inertial_quat)
        return out


def link_aabb_inertia(link: LinkDesc) -> np.ndarray:
    """Bullet's compound-AABB box inertia in the link's inertial frame (see module docstring)."""
    if link.inertia_diag is not None:
        return np.asarray(link.inertia_diag, dtype=np.float64)
    if link.mass <= 0 or not link.shapes:
        return np.zeros(3)
    ip, iq = X.tf_inv(link.inertial_pos, link.inertial_quat)
    lo = np.full(3, np.inf); hi = np.full(3, -np.inf)
    for s in link.shapes:
        p, r = X.tf_mul(ip, iq, s.pos, s.quat)          # shape frame in the inertial frame
        R = X.quat_to_mat(r)
        smin, smax = s.local_aabb()
        c = 0.5 * (smin + smax); h = 0.5 * (smax - smin)
        cw = R @ c + p
        hw = np.abs(R) @ h
        lo = np.minimum(lo, cw - hw); hi = np.maximum(hi, cw + hw)
    l = hi - lo
    return link.mass / 12.0 * np.array([l[1] ** 2 + l[2] ** 2, l[0] ** 2 + l[2] ** 2, l[0] ** 2 + l[1] ** 2])


def link_contact_threshold(link: LinkDesc) -> float:
    """`btCollisionShape::getContactBreakingThreshold(0.02)` of the link's compound shape: 0.02 × (AABB
    half-diagonal + |AABB centre|), all in the inertial frame [UPSTREAM-BULLET, unverified]."""
    if not link.shapes:
        return 0.0
    ip, iq = X.tf_inv(link.inertial_pos, link.inertial_quat)
    lo = np.full(3, np.inf); hi = np.full(3, -np.inf)
    for s in link.shapes:
        p, r = X.tf_mul(ip, iq, s.pos, s.quat)
        R = X.quat_to_mat(r)
        smin, smax = s.local_aabb()
        c = 0.5 * (smin + smax); h = 0.5 * (smax - smin)
        cw = R @ c + p
        hw = np.abs(R) @ h
        lo = np.minimum(lo, cw - hw); hi = np.maximum(hi, cw + hw)
    centre = 0.5 * (lo + hi)
    return 0.02 * (0.5 * np.linalg.norm(hi - lo) + np.linalg.norm(centre))


# ---------------------------------------------------------------------------------------------------------------
@dataclass
class DynBody:
    """One integrated body: frame at the composite COM, principal axes."""
    art: int                        # articulation id (index of the source MultiBodyDesc in the scene)
    parent: int                     # index into the dyn-body list, -1 = static world
    jtype: int                      # JOINT_REVOLUTE / JOINT_PRISMATIC / JOINT_FREE
    ta_pos: np.ndarray              # parent body frame (or world) -> joint frame at q=0
    ta_quat: np.ndarray
    axis: np.ndarray                # joint axis in the joint frame (unit)
    tb_pos: np.ndarray              # joint frame (after the joint motion) -> body frame
    tb_quat: np.ndarray
    mass: float
    inertia: np.ndarray             # diag, body frame
    gravity: np.ndarray
    ref_joint: int                  # PyBullet joint index of the driving joint (-1 for free bodies)
    lower: float = 0.0
    upper: float = -1.0
    limit_enforced: bool = False
    root_link: int = -1             # PyBullet link index of the composite's root link
    init_pos: Optional[np.ndarray] = None   # free bodies: world pose of the body frame at reset
    init_quat: Optional[np.ndarray] = None
    damping: float = 0.0            # joint damping torque -damping * qd (URDF <dynamics damping>)


@dataclass
class Attached:
    """Something rigidly attached to a dyn body (or the static world when body == -1)."""
    body: int
    pos: np.ndarray
    quat: np.ndarray


def _rot_inertia(R: np.ndarray, diag: np.ndarray) -> np.ndarray:
    return R @ np.diag(diag) @ R.T


def reduce_bodies(mb: MultiBodyDesc, art: int, q_reset: Dict[int, float], frozen: set,
                  body_offset: int) -> Tuple[List[DynBody], Dict[int, Attached]]:
    """Reduce `mb` to dynamic bodies (see module docstring).

    q_reset: joint values used for frozen joints (and to place free bodies). frozen: PyBullet joint indices that the
    reference freezes. Returns (bodies, link_attach) where link_attach[ref_link] gives the LINK frame of every
    PyBullet link relative to its dyn body (or the world for static links).
    """
    n = len(mb.links)
    # 1. group links into composites
    movable = {}
    for l in mb.links:
        movable[l.ref_index] = (l.jtype in ("revolute", "prismatic")) and (l.ref_index not in frozen)
    base_dynamic = not mb.fixed_base
    comp_root: Dict[int, int] = {}      # link -> root link of its composite (-2 = static world)
    comp_root[-1] = -1 if base_dynamic else -2
    for l in mb.links:                  # parents precede children in PyBullet order
        if movable[l.ref_index]:
            comp_root[l.ref_index] = l.ref_index
        else:
            comp_root[l.ref_index] = comp_root[l.parent]
    roots = sorted({r for r in comp_root.values() if r != -2})

    # 2. link frames relative to their composite root's LINK frame (joint values frozen at q_reset)
    rel: Dict[int, Tuple[np.ndarray, np.ndarray]] = {}
    world_frames = mb.link_frames(q_reset)
    for i in [-1] + list(range(n)):
        r = comp_root[i]
        if r == -2:
            rel[i] = world_frames[i]            # static: absolute world pose
        elif r == i:
            rel[i] = (np.zeros(3), np.array([0.0, 0, 0, 1]))
        else:
            l = mb.link(i)
            pp, pq = rel[l.parent]
            p, rq = X.tf_mul(pp, pq, l.pos, l.quat)
            v = q_reset.get(i, 0.0)
            if l.jtype == "revolute":
                rq = X.quat_normalize(X.quat_mul(rq, X.quat_from_axis_angle(l.axis, v)))
            elif l.jtype == "prismatic":
                p = p + X.quat_rotate(rq, l.axis * v)
            rel[i] = (p, rq)

    # 3. composite inertia per root, expressed in the root LINK frame
    body_frame: Dict[int, Tuple[np.ndarray, np.ndarray]] = {}   # root -> body frame relative to the root LINK frame
    body_mass: Dict[int, float] = {}
    body_inertia: Dict[int, np.ndarray] = {}
    for r in roots:
        members = [i for i in [-1] + list(range(n)) if comp_root[i] == r]
        m_tot = 0.0; mc = np.zeros(3)
        parts = []
        for i in members:
            l = mb.link(i)
            if l.mass <= 0:
                continue
            p, rq = rel[i]
            cp, cq = X.tf_mul(p, rq, l.inertial_pos, l.inertial_quat)
            Ii = _rot_inertia(X.quat_to_mat(cq), link_aabb_inertia(l))
            parts.append((l.mass, cp, Ii))
            m_tot += l.mass; mc += l.mass * cp
        if m_tot > 0:
            com = mc / m_tot
            I = np.zeros((3, 3))
            for m, cp, Ii in parts:
                d = cp - com
                I += Ii + m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
            w, V = np.linalg.eigh(I)
            if len(parts) == 1:                  # keep the inertial frame of a single link (already diagonal)
                l0 = [mb.link(i) for i in members if mb.link(i).mass > 0][0]
                i0 = [i for i in members if mb.link(i).mass > 0][0]
                p, rq = rel[i0]
                cp, cq = X.tf_mul(p, rq, l0.inertial_pos, l0.inertial_quat)
                body_frame[r] = (cp, cq)
                body_inertia[r] = link_aabb_inertia(l0)
            else:
                if np.linalg.det(V) < 0:
                    V[:, 2] = -V[:, 2]
                body_frame[r] = (com, X.mat_to_quat(V))
                body_inertia[r] = w
            body_mass[r] = m_tot
        else:
            l = mb.link(r)
            body_frame[r] = (l.inertial_pos.copy(), l.inertial_quat.copy())
            body_mass[r] = 0.0
            body_inertia[r] = np.zeros(3)

    # 4. dyn bodies
    index_of = {r: body_offset + k for k, r in enumerate(roots)}
    bodies: List[DynBody] = []
    for r in roots:
        bf_p, bf_q = body_frame[r]
        if r == -1:     # free base
            wp, wq = X.tf_mul(mb.base_pos, mb.base_quat, bf_p, bf_q)
            bodies.append(DynBody(art=art, parent=-1, jtype=JOINT_FREE, ta_pos=np.zeros(3), ta_quat=np.array([0.0, 0, 0, 1]),
                                  axis=np.array([0.0, 0, 1]), tb_pos=np.zeros(3), tb_quat=np.array([0.0, 0, 0, 1]),
                                  mass=body_mass[r], inertia=body_inertia[r], gravity=mb.gravity.copy(), ref_joint=-1,
                                  root_link=-1, init_pos=wp, init_quat=wq))
            continue
        l = mb.link(r)
        pr = comp_root[l.parent]
        # joint frame at q=0 relative to the parent LINK frame chain: rel[parent] ∘ (l.pos, l.quat)
        jp, jq = X.tf_mul(*rel[l.parent], l.pos, l.quat)
        if pr == -2:
            ta_p, ta_q = jp, jq                  # absolute world pose
            parent_idx = -1
        else:
            ip, iq = X.tf_inv(*body_frame[pr])
            ta_p, ta_q = X.tf_mul(ip, iq, jp, jq)
            parent_idx = index_of[pr]
        bodies.append(DynBody(art=art, parent=parent_idx,
                              jtype=JOINT_REVOLUTE if l.jtype == "revolute" else JOINT_PRISMATIC,
                              ta_pos=ta_p, ta_quat=ta_q, axis=l.axis / np.linalg.norm(l.axis),
                              tb_pos=bf_p, tb_quat=bf_q, mass=body_mass[r], inertia=body_inertia[r],
                              gravity=mb.gravity.copy(), ref_joint=r, lower=l.lower, upper=l.upper,
                              limit_enforced=l.limit_enforced, root_link=r, damping=l.damping))

    # 5. attachment of every PyBullet LINK frame
    attach: Dict[int, Attached] = {}
    for i in [-1] + list(range(n)):
        r = comp_root[i]
        if r == -2:
            attach[i] = Attached(-1, *rel[i])
        else:
            ip, iq = X.tf_inv(*body_frame[r])
            p, rq = X.tf_mul(ip, iq, *rel[i])
            attach[i] = Attached(index_of[r], p, rq)
    return bodies, attach
