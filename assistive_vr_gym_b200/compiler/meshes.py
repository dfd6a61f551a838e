"""Mesh loaders and convex-hull preparation for the model compiler (CPU, offline).

The reference hands mesh files to PyBullet (`loadURDF` / `createCollisionShape(GEOM_MESH)`,
reference `world_creation.py:274-293`, `human_creation.py:97-99`), which turns every `<collision><mesh>` into a
convex hull of the file's vertices (one hull per `o` group for OBJ files = VHACD compounds).
This module restates that behaviour: it parses OBJ / binary STL / COLLADA, applies the COLLADA up-axis and unit
conversion, and returns hull vertices + face planes ready for the device-side support-function narrowphase.
"""
from __future__ import annotations

import re
import struct
import xml.etree.ElementTree as ET
from typing import List, Tuple

import numpy as np
from scipy.spatial import ConvexHull

MAX_HULL_VERTS = 48   # device hull support loops are bounded by this (see csrc/avg_kernels.cu)
MAX_HULL_FACES = 64


def load_obj_groups(path: str) -> List[np.ndarray]:
    """Vertices per `o` group of a Wavefront OBJ (one convex hull per group)."""
    groups: List[List[List[float]]] = []
    cur: List[List[float]] | None = None
    with open(path, "r") as f:
        for line in f:
            if line.startswith("o "):
                cur = []
                groups.append(cur)
            elif line.startswith("v "):
                if cur is None:
                    cur = []
                    groups.append(cur)
                parts = line.split()
                cur.append([float(parts[1]), float(parts[2]), float(parts[3])])
    return [np.asarray(g, dtype=np.float64) for g in groups if len(g) >= 4]


def load_stl(path: str) -> np.ndarray:
    """All vertices of a binary (or ASCII) STL file."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:5].lower() == b"solid" and b"facet" in data[:400]:
        vs = re.findall(rb"vertex\s+(\S+)\s+(\S+)\s+(\S+)", data)
        return np.asarray([[float(a), float(b), float(c)] for a, b, c in vs], dtype=np.float64)
    ntri = struct.unpack_from("<I", data, 80)[0]
    rec = np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("attr", "<u2")])
    tris = np.frombuffer(data, dtype=rec, count=ntri, offset=84)
    return tris["v"].reshape(-1, 3).astype(np.float64)


def _node_matrix(node, ns) -> np.ndarray:
    m = np.eye(4)
    for child in node:
        tag = child.tag.split("}")[-1]
        vals = [float(x) for x in (child.text or "").split()]
        if tag == "matrix" and len(vals) == 16:
            m = m @ np.asarray(vals).reshape(4, 4)
        elif tag == "translate" and len(vals) == 3:
            t = np.eye(4); t[:3, 3] = vals; m = m @ t
        elif tag == "scale" and len(vals) == 3:
            m = m @ np.diag(vals + [1.0])
        elif tag == "rotate" and len(vals) == 4:
            ax = np.asarray(vals[:3]); ang = np.deg2rad(vals[3])
            n = np.linalg.norm(ax)
            if n > 0:
                ax = ax / n
                K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
                R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
                t = np.eye(4); t[:3, :3] = R; m = m @ t
    return m


def load_dae(path: str, apply_up_axis: bool = False) -> np.ndarray:
    """Vertices of every instanced geometry in a COLLADA file, in metres.

    Mirrors what Bullet's URDF importer does for `<collision><mesh>` COLLADA files [UPSTREAM-BULLET, unverified]:
    node transforms are applied to their instanced geometry and `<unit meter=…>` scales, but the `<up_axis>`
    rotation is NOT applied (URDF meshes are authored in the link frame regardless of the tag). This is confirmed
    by the assets themselves: every `.dae` in `assets/jaco/meshes` — including the `Y_UP` ones — has the same
    bounding box as its `.STL` twin only without the rotation, and only then do the link geometries line up with the
    joint origins of `j2s7s300_gym.urdf` (checked in `tests/test_compiler.py`).
    """
    tree = ET.parse(path)
    root = tree.getroot()
    ns = {"c": root.tag.split("}")[0].strip("{")} if "}" in root.tag else {}
    pre = "c:" if ns else ""

    def find(el, p):
        return el.find(p, ns)

    def findall(el, p):
        return el.findall(p, ns)

    unit = 1.0
    up = "Y_UP"
    asset = find(root, f"{pre}asset")
    if asset is not None:
        u = find(asset, f"{pre}unit")
        if u is not None and u.get("meter"):
            unit = float(u.get("meter"))
        ua = find(asset, f"{pre}up_axis")
        if ua is not None and ua.text:
            up = ua.text.strip()

    geoms = {}
    for g in root.iter(("{%s}geometry" % ns["c"]) if ns else "geometry"):
        gid = g.get("id")
        mesh = find(g, f"{pre}mesh")
        if mesh is None:
            continue
        sources = {}
        for s in findall(mesh, f"{pre}source"):
            fa = find(s, f"{pre}float_array")
            if fa is not None and fa.text:
                sources[s.get("id")] = np.asarray(fa.text.split(), dtype=np.float64)
        verts_el = find(mesh, f"{pre}vertices")
        pos = None
        if verts_el is not None:
            for inp in findall(verts_el, f"{pre}input"):
                if inp.get("semantic") == "POSITION":
                    pos = sources.get(inp.get("source").lstrip("#"))
        if pos is None:
            continue
        geoms[gid] = pos.reshape(-1, 3)

    out = []

    def walk(node, m):
        m = m @ _node_matrix(node, ns)
        for ig in findall(node, f"{pre}instance_geometry"):
            gid = ig.get("url", "").lstrip("#")
            if gid in geoms:
                v = geoms[gid]
                out.append(v @ m[:3, :3].T + m[:3, 3])
        for ch in findall(node, f"{pre}node"):
            walk(ch, m)

    for vs in root.iter(("{%s}visual_scene" % ns["c"]) if ns else "visual_scene"):
        for node in findall(vs, f"{pre}node"):
            walk(node, np.eye(4))
    if not out:  # no scene graph: take raw geometry
        out = list(geoms.values())
    v = np.concatenate(out, axis=0) * unit
    if not apply_up_axis:
        return v
    if up == "Y_UP":     # +90° about X: (x, y, z) -> (x, -z, y)
        v = np.stack([v[:, 0], -v[:, 2], v[:, 1]], axis=1)
    elif up == "X_UP":   # +90° about Y: x -> z
        v = np.stack([-v[:, 2], v[:, 1], v[:, 0]], axis=1)
    return v


def load_mesh_hulls(path: str) -> List[np.ndarray]:
    """Vertex sets, one per convex piece, for a collision mesh file."""
    low = path.lower()
    if low.endswith(".obj"):
        return load_obj_groups(path)
    if low.endswith(".stl"):
        return [load_stl(path)]
    if low.endswith(".dae"):
        return [load_dae(path)]
    raise ValueError(f"unsupported mesh format: {path}")


def _hull_planes(points: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Unique outward face planes (n, d) with n·x <= d inside, merged over coplanar triangles."""
    hull = ConvexHull(points)
    eq = hull.equations  # n·x + off <= 0 inside
    planes = []
    for n0, n1, n2, off in eq:
        n = np.array([n0, n1, n2]); d = -off
        dup = False
        for pn, pd in planes:
            if np.dot(pn, n) > 1.0 - 1e-6 and abs(pd - d) < 1e-7:
                dup = True
                break
        if not dup:
            planes.append((n, d))
    return hull.vertices, np.asarray([np.append(n, d) for n, d in planes])


def simplify_hull(points: np.ndarray, max_verts: int = MAX_HULL_VERTS) -> np.ndarray:
    """Convex-hull vertices of `points`, reduced to at most `max_verts`.

    Bullet keeps every hull vertex; the device narrowphase bounds its support loop, so hulls with more vertices are
    reduced greedily: start from the 6 axis-extreme vertices and repeatedly add the hull vertex farthest outside
    the current inner hull. The result is an inner approximation; the worst-case shrink is reported by the compiler
    (`scene.py` prints it and DESIGN.md records the numbers).
    """
    pts = np.unique(np.round(points, 9), axis=0)
    hull = ConvexHull(pts)
    hv = pts[hull.vertices]
    if len(hv) <= max_verts:
        return hv
    chosen = set()
    for ax in range(3):
        chosen.add(int(np.argmin(hv[:, ax]))); chosen.add(int(np.argmax(hv[:, ax])))
    # make sure the seed is non-degenerate
    dirs = np.array([[1, 1, 1], [1, -1, 1], [-1, 1, 1], [-1, -1, 1]], dtype=float)
    for d in dirs:
        chosen.add(int(np.argmax(hv @ d))); chosen.add(int(np.argmin(hv @ d)))
    chosen = sorted(chosen)
    while len(chosen) < max_verts:
        sub = hv[chosen]
        h = ConvexHull(sub)
        # distance of every candidate outside the current hull
        dist = np.max(hv @ h.equations[:, :3].T + h.equations[:, 3], axis=1)
        dist[chosen] = -1.0
        k = int(np.argmax(dist))
        if dist[k] <= 1e-9:
            break
        chosen.append(k)
    sub = hv[sorted(chosen)]
    h = ConvexHull(sub)
    return sub[h.vertices]


def hull_error(points: np.ndarray, hull_verts: np.ndarray) -> float:
    """Largest distance of an original point outside the reduced hull (metres)."""
    h = ConvexHull(hull_verts)
    d = np.max(points @ h.equations[:, :3].T + h.equations[:, 3], axis=1)
    return float(max(0.0, d.max()))


def prepare_hull(points: np.ndarray, max_verts: int = MAX_HULL_VERTS,
                 max_faces: int = MAX_HULL_FACES) -> Tuple[np.ndarray, np.ndarray, float]:
    """-> (vertices[k,3], planes[f,4] (n,d with n·x<=d inside), shrink error in metres)."""
    verts = simplify_hull(points, max_verts)
    # face planes, merged; if there are still too many faces keep the largest-area ones (planes are only used by
    # the deep-penetration fallback)
    h = ConvexHull(verts)
    eq = h.equations
    areas = {}
    keys = []
    for simplex, (n0, n1, n2, off) in zip(h.simplices, eq):
        a, b, c = verts[simplex]
        area = 0.5 * np.linalg.norm(np.cross(b - a, c - a))
        key = None
        for k in keys:
            if np.dot(k[:3], [n0, n1, n2]) > 1.0 - 1e-6 and abs(k[3] + off) < 1e-7:
                key = k
                break
        if key is None:
            key = (n0, n1, n2, -off)
            keys.append(key)
            areas[key] = 0.0
        areas[key] += area
    keys.sort(key=lambda k: -areas[k])
    planes = np.asarray(keys[:max_faces], dtype=np.float64)
    err = hull_error(points, verts)
    return verts, planes, err
