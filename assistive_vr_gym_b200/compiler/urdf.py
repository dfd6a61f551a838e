"""Minimal URDF reader producing PyBullet's link numbering.

PyBullet numbers links/joints in depth-first order of the URDF tree with children visited in file order
(SURVEY.md §2.1 re-derived every hard-coded index of the reference this way, e.g. Jaco arm joints 1..7, end effector
link 8, finger joints 9/11/13 used at reference `world_creation.py:283,320,332`). `<gazebo>`, `<transmission>`,
`<mimic>`, `<safety_controller>` and XML comments are ignored, as in `loadURDF`.
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import xform as X


@dataclass
class Geometry:
    kind: str                      # 'box' | 'sphere' | 'cylinder' | 'capsule' | 'mesh'
    size: np.ndarray = field(default_factory=lambda: np.zeros(3))   # box: full extents
    radius: float = 0.0
    length: float = 0.0
    filename: str = ""
    scale: np.ndarray = field(default_factory=lambda: np.ones(3))


@dataclass
class Collision:
    pos: np.ndarray
    quat: np.ndarray
    geom: Geometry


@dataclass
class Link:
    name: str
    mass: float = 0.0
    inertial_pos: np.ndarray = field(default_factory=lambda: np.zeros(3))
    inertial_quat: np.ndarray = field(default_factory=lambda: np.array([0.0, 0, 0, 1]))
    inertia_diag_file: np.ndarray = field(default_factory=lambda: np.zeros(3))
    inertia_file: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))   # full <inertia> tensor, inertial frame
    collisions: List[Collision] = field(default_factory=list)
    lateral_friction: float = 0.5  # Bullet default


@dataclass
class Joint:
    name: str
    jtype: str                     # 'fixed' | 'revolute' | 'continuous' | 'prismatic'
    parent: str
    child: str
    pos: np.ndarray
    quat: np.ndarray
    axis: np.ndarray
    lower: float = 0.0
    upper: float = -1.0            # PyBullet reports (0, -1) when no limit is given
    damping: float = 0.0
    has_limit_tag: bool = False


@dataclass
class UrdfModel:
    name: str
    base: str
    links: Dict[str, Link]
    joints: List[Joint]
    # PyBullet ordering: index i is joint i and its child link i
    order: List[Joint] = field(default_factory=list)
    link_index: Dict[str, int] = field(default_factory=dict)   # base -> -1

    def parent_index(self, i: int) -> int:
        return self.link_index[self.order[i].parent]


def _vec(s: Optional[str], default):
    if s is None:
        return np.asarray(default, dtype=np.float64)
    return np.asarray([float(x) for x in s.split()], dtype=np.float64)


def _origin(el):
    if el is None:
        return np.zeros(3), np.array([0.0, 0, 0, 1])
    return _vec(el.get("xyz"), [0, 0, 0]), X.quat_from_euler(_vec(el.get("rpy"), [0, 0, 0]))


def parse_urdf(path: str) -> UrdfModel:
    root = ET.parse(path).getroot()
    base_dir = os.path.dirname(path)
    links: Dict[str, Link] = {}
    for le in root.findall("link"):
        link = Link(name=le.get("name"))
        inert = le.find("inertial")
        if inert is not None:
            m = inert.find("mass")
            if m is not None:
                link.mass = float(m.get("value"))
            link.inertial_pos, link.inertial_quat = _origin(inert.find("origin"))
            ie = inert.find("inertia")
            if ie is not None:
                link.inertia_diag_file = np.array([float(ie.get("ixx", 0)), float(ie.get("iyy", 0)), float(ie.get("izz", 0))])
                ixy, ixz, iyz = float(ie.get("ixy", 0)), float(ie.get("ixz", 0)), float(ie.get("iyz", 0))
                d = link.inertia_diag_file
                link.inertia_file = np.array([[d[0], ixy, ixz], [ixy, d[1], iyz], [ixz, iyz, d[2]]])
        contact = le.find("contact")
        if contact is not None:
            lf = contact.find("lateral_friction")
            if lf is not None:
                link.lateral_friction = float(lf.get("value"))
        for ce in le.findall("collision"):
            pos, quat = _origin(ce.find("origin"))
            ge = ce.find("geometry")
            if ge is None:
                continue
            g = None
            if ge.find("box") is not None:
                g = Geometry("box", size=_vec(ge.find("box").get("size"), [0, 0, 0]))
            elif ge.find("sphere") is not None:
                g = Geometry("sphere", radius=float(ge.find("sphere").get("radius")))
            elif ge.find("cylinder") is not None:
                c = ge.find("cylinder")
                g = Geometry("cylinder", radius=float(c.get("radius")), length=float(c.get("length")))
            elif ge.find("capsule") is not None:
                c = ge.find("capsule")
                g = Geometry("capsule", radius=float(c.get("radius")), length=float(c.get("length")))
            elif ge.find("mesh") is not None:
                me = ge.find("mesh")
                fn = me.get("filename")
                if fn.startswith("package://"):
                    fn = fn[len("package://"):]
                g = Geometry("mesh", filename=os.path.normpath(os.path.join(base_dir, fn)),
                             scale=_vec(me.get("scale"), [1, 1, 1]))
            if g is not None:
                link.collisions.append(Collision(pos, quat, g))
        links[link.name] = link

    joints: List[Joint] = []
    for je in root.findall("joint"):
        if je.find("parent") is None or je.find("child") is None:
            continue  # <transmission><joint> style entries live elsewhere, but be safe
        pos, quat = _origin(je.find("origin"))
        axis = _vec(je.find("axis").get("xyz") if je.find("axis") is not None else None, [1, 0, 0])
        j = Joint(name=je.get("name"), jtype=je.get("type"), parent=je.find("parent").get("link"),
                  child=je.find("child").get("link"), pos=pos, quat=quat, axis=axis)
        lim = je.find("limit")
        if lim is not None and (lim.get("lower") is not None or lim.get("upper") is not None):
            j.lower = float(lim.get("lower", 0.0)); j.upper = float(lim.get("upper", 0.0)); j.has_limit_tag = True
        dyn = je.find("dynamics")
        if dyn is not None:
            j.damping = float(dyn.get("damping", 0.0))
        joints.append(j)

    children = {j.child for j in joints}
    bases = [n for n in links if n not in children]
    base = bases[0]
    model = UrdfModel(name=root.get("name", ""), base=base, links=links, joints=joints)
    model.link_index[base] = -1

    def dfs(parent_name: str):
        for j in joints:               # file order
            if j.parent == parent_name:
                model.link_index[j.child] = len(model.order)
                model.order.append(j)
                dfs(j.child)

    dfs(base)
    return model
