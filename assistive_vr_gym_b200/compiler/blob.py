"""ModelBlob writer/reader: numpy mirrors of the structs in `include/avg_model.h`."""
from __future__ import annotations

import numpy as np

from . import xform as X
from .mbody import JOINT_FREE, SHAPE_HULL, SHAPE_COMPOUND
from .scene import CompiledScene, _world_aabb

AVG_MAGIC = 0x4D475641
AVG_VERSION = 10
ENV_STRIDE = 192

BODY_DT = np.dtype([
    ("parent", "<i4"), ("jtype", "<i4"), ("dof", "<i4"), ("qidx", "<i4"),
    ("ta_pos", "<f4", 3), ("ta_quat", "<f4", 4), ("axis", "<f4", 3), ("tb_pos", "<f4", 3), ("tb_quat", "<f4", 4),
    ("mass", "<f4"), ("inertia", "<f4", 3), ("gravity", "<f4", 3), ("anc_mask", "<u4"),
    ("ref_body", "<i4"), ("ref_joint", "<i4"), ("sub_end", "<i4"),
])
DOF_DT = np.dtype([
    ("body", "<i4"), ("flags", "<u4"), ("lower", "<f4"), ("upper", "<f4"), ("rep_lower", "<f4"), ("rep_upper", "<f4"),
    ("kp", "<f4"), ("kd", "<f4"), ("max_force", "<f4"), ("action", "<i4"), ("human_slot", "<i4"), ("init_target", "<f4"),
    ("damping", "<f4"), ("pad", "<i4", 3),
])
SHAPE_DT = np.dtype([
    ("type", "<i4"), ("body", "<i4"), ("ref_body", "<i4"), ("ref_link", "<i4"),
    ("pos", "<f4", 3), ("quat", "<f4", 4), ("radius", "<f4"), ("half", "<f4", 3), ("margin", "<f4"),
    ("vert_off", "<i4"), ("vert_cnt", "<i4"), ("plane_off", "<i4"), ("plane_cnt", "<i4"),
    ("friction", "<f4"), ("thr", "<f4"), ("aabb_c", "<f4", 3), ("aabb_h", "<f4", 3), ("pad", "<i4", 4),
])
FRAME_DT = np.dtype([("body", "<i4"), ("pos", "<f4", 3), ("quat", "<f4", 4)])
BPS_DT = np.dtype([("c", "<f4", 3), ("h", "<f4", 3), ("thr", "<f4"), ("mask", "<u4")])
HEADER_DT = np.dtype([
    ("magic", "<u4"), ("version", "<u4"), ("total_bytes", "<u4"), ("task", "<i4"),
    ("n_body", "<i4"), ("n_ebody", "<i4"), ("n_dof", "<i4"), ("n_jdof", "<i4"), ("n_free", "<i4"),
    ("n_shape", "<i4"), ("n_mshape", "<i4"), ("n_vert", "<i4"), ("n_plane", "<i4"), ("n_pair", "<i4"), ("n_frame", "<i4"),
    ("substeps", "<i4"), ("solver_iters", "<i4"),
    ("n_action_robot", "<i4"), ("n_action_human", "<i4"), ("n_obs_robot", "<i4"), ("n_obs_human", "<i4"),
    ("human_control", "<i4"),
    ("dt", "<f4"), ("erp", "<f4"), ("lin_damp", "<f4"), ("ang_damp", "<f4"), ("residual_thr", "<f4"), ("max_vel", "<f4"),
    ("action_scale", "<f4"), ("weld_max_force", "<f4"), ("weld_body_a", "<i4"), ("weld_body_b", "<i4"),
    ("task_f", "<f4", 32),
    ("off_body", "<u4"), ("off_dof", "<u4"), ("off_shape", "<u4"), ("off_vert", "<u4"), ("off_plane", "<u4"),
    ("off_pair", "<u4"), ("off_frame", "<u4"), ("off_bps", "<u4"), ("off_bpm", "<u4"),
    ("n_block", "<i4"), ("block_start", "<i4", 4), ("off_bcap", "<u4"),
    ("off_mlp", "<u4"), ("n_mlp", "<i4"), ("mlp_dof", "<i4", 4),
    ("off_target", "<u4"), ("n_target", "<i4"), ("n_target_upper", "<i4"),
    ("n_cshape", "<i4"), ("off_caabb", "<u4"), ("n_internal", "<i4"), ("n_particle", "<i4"), ("pshape", "<i4"),
    ("p_mass", "<f4"), ("p_gravity", "<f4", 3), ("tool_body", "<i4"), ("head_frozen_mask", "<i4"), ("warmstart", "<f4"), ("pad2", "<u4", 4),
])
assert BODY_DT.itemsize == 128 and DOF_DT.itemsize == 64 and SHAPE_DT.itemsize == 128 and FRAME_DT.itemsize == 32


def bounding_capsule(d) -> tuple:
    """(p0, p1, r) in the shape frame: a capsule that contains the shape including its collision margin.  Used by the
    device narrowphase as a cheap conservative cull before GJK (elongated robot links have fat AABBs)."""
    from .mbody import SHAPE_SPHERE, SHAPE_CAPSULE, SHAPE_BOX, SHAPE_CYLINDER, SHAPE_HULL
    z = np.zeros(3)
    if d.kind == SHAPE_SPHERE:
        return z, z, float(d.radius)
    if d.kind in (SHAPE_CAPSULE, SHAPE_CYLINDER):
        hl = float(d.half[2])
        return np.array([0, 0, -hl]), np.array([0, 0, hl]), float(d.radius)
    if d.kind == SHAPE_BOX:
        k = int(np.argmax(d.half))
        a = np.zeros(3); a[k] = d.half[k]
        others = [d.half[i] for i in range(3) if i != k]
        return -a, a, float(np.hypot(*others))
    if d.kind in (SHAPE_HULL, SHAPE_COMPOUND):
        if d.kind == SHAPE_COMPOUND:
            v = np.concatenate([np.asarray(c.verts, dtype=np.float64) @ X.quat_to_mat(c.quat).T + c.pos for c in d.children], axis=0)
        else:
            v = np.asarray(d.verts, dtype=np.float64)
        c = v.mean(0)
        w, V = np.linalg.eigh(np.cov((v - c).T))
        a = V[:, -1]
        t = (v - c) @ a
        perp = np.linalg.norm((v - c) - np.outer(t, a), axis=1)
        return c + a * t.min(), c + a * t.max(), float(perp.max() + 0.001)
    return z, z, 1e9                     # plane: never culled


def _align(n: int, a: int = 16) -> int:
    return (n + a - 1) // a * a


def scene_to_blob(scene: CompiledScene, overrides: dict | None = None) -> bytes:
    nb = len(scene.bodies)
    bodies = np.zeros(nb, dtype=BODY_DT)
    for i, b in enumerate(scene.bodies):
        r = bodies[i]
        r["parent"] = b.parent; r["jtype"] = b.jtype; r["dof"] = b.dof; r["qidx"] = b.qidx
        r["ta_pos"] = b.ta_pos; r["ta_quat"] = b.ta_quat; r["axis"] = b.axis
        r["tb_pos"] = b.tb_pos; r["tb_quat"] = b.tb_quat
        r["mass"] = b.mass; r["inertia"] = b.inertia; r["gravity"] = b.gravity
        mask = 1 << i
        p = b.parent
        while p >= 0:
            mask |= 1 << p
            p = scene.bodies[p].parent
        r["anc_mask"] = mask
        r["ref_body"] = scene.multibodies[b.art].ref_body; r["ref_joint"] = b.ref_joint
    # depth-first order: the subtree of body i is the contiguous range [i, sub_end) (the dynamics kernel sums over it)
    for i in range(nb):
        sub = [j for j in range(nb) if (int(bodies[j]["anc_mask"]) >> i) & 1]
        assert sub == list(range(i, i + len(sub))), f"bodies are not in depth-first order at body {i}: {sub}"
        bodies[i]["sub_end"] = i + len(sub)
    dofs = np.zeros(len(scene.dofs), dtype=DOF_DT)
    for i, d in enumerate(scene.dofs):
        for k in ("body", "flags", "lower", "upper", "rep_lower", "rep_upper", "kp", "kd", "max_force", "action",
                  "human_slot", "init_target"):
            dofs[i][k] = d[k]
        dofs[i]["damping"] = d.get("damping", 0.0)
    verts = []
    planes = []
    cshapes = list(getattr(scene, "cshapes", None) or [])
    particle = getattr(scene, "particle", None)
    all_shapes = list(scene.shapes) + cshapes
    n_top = len(scene.shapes)
    shapes = np.zeros(len(all_shapes) + (1 if particle is not None else 0), dtype=SHAPE_DT)
    hull_index = {}
    first_child = {}
    for k, c in enumerate(cshapes):
        first_child.setdefault(c.parent, n_top + k)
    for i, s in enumerate(all_shapes):
        r = shapes[i]
        d = s.desc
        r["type"] = d.kind; r["body"] = s.body; r["ref_body"] = s.ref_body; r["ref_link"] = s.ref_link
        r["pos"] = s.pos; r["quat"] = s.quat; r["radius"] = d.radius; r["half"] = d.half; r["margin"] = s.margin
        if d.kind == SHAPE_HULL:
            key = id(d.verts)
            if key not in hull_index:
                hull_index[key] = (sum(len(v) for v in verts), len(d.verts), sum(len(p) for p in planes), len(d.planes))
                verts.append(np.asarray(d.verts, dtype=np.float32))
                planes.append(np.asarray(d.planes, dtype=np.float32))
            r["vert_off"], r["vert_cnt"], r["plane_off"], r["plane_cnt"] = hull_index[key]
        if d.kind == SHAPE_COMPOUND:
            r["vert_off"] = first_child[i]; r["vert_cnt"] = len(s.children)
        r["friction"] = d.friction; r["thr"] = s.thr
        r["pad"][0] = s.parent if i >= n_top else -1          # children: their compound
        if s.body >= 0:
            lo, hi = d.local_aabb()
            r["aabb_c"] = 0.5 * (lo + hi); r["aabb_h"] = 0.5 * (hi - lo)
        else:
            c, h = _world_aabb(d, s.pos, s.quat)
            r["aabb_c"] = c; r["aabb_h"] = h
    if particle is not None:                                  # template of the food / water spheres (pose from the particle record)
        r = shapes[len(all_shapes)]
        r["type"] = particle.kind; r["body"] = -1; r["ref_body"] = 7; r["ref_link"] = -1; r["quat"] = [0, 0, 0, 1]
        r["radius"] = particle.radius; r["margin"] = particle.radius; r["friction"] = particle.friction; r["thr"] = scene.particle_thr
        r["aabb_h"] = [particle.radius] * 3; r["pad"][0] = -1
    # child AABBs in the frame of the owning body (the device tests them against points / boxes brought into that frame)
    caabb = np.zeros((len(cshapes), 8), dtype="<f4")
    for k, c in enumerate(cshapes):
        R = X.quat_to_mat(c.quat)
        lo, hi = c.desc.local_aabb()
        caabb[k, 0:3] = R @ (0.5 * (lo + hi)) + c.pos; caabb[k, 4:7] = np.abs(R) @ (0.5 * (hi - lo))
    verts = np.concatenate(verts, axis=0).astype("<f4") if verts else np.zeros((0, 3), "<f4")
    verts = np.concatenate([verts, np.zeros((len(verts), 1), "<f4")], axis=1)          # padded to float4 for 16-byte loads
    planes = np.concatenate(planes, axis=0).astype("<f4") if planes else np.zeros((0, 4), "<f4")
    opairs = scene.opairs if getattr(scene, "opairs", None) is not None else scene.pairs
    assert len(shapes) < 32768
    pairs = (opairs[:, 0].astype("<u4") | (opairs[:, 1].astype("<u4") << 16)).astype("<u4")
    frames = np.zeros(len(scene.frames), dtype=FRAME_DT)
    for i, (b, p, q) in enumerate(scene.frames):
        frames[i]["body"] = b; frames[i]["pos"] = p; frames[i]["quat"] = q

    # broadphase tables (same pair set as `pairs`, organised for the device loop)
    nms = scene.n_mshape
    bps = np.zeros(len(scene.shapes) - nms, dtype=BPS_DT)
    bpm = np.zeros(nms, dtype="<u4")
    for i in range(nms, len(scene.shapes)):
        bps[i - nms]["c"] = shapes[i]["aabb_c"]; bps[i - nms]["h"] = shapes[i]["aabb_h"]; bps[i - nms]["thr"] = shapes[i]["thr"]
    for a, b in scene.pairs:
        if b >= nms:
            bps[b - nms]["mask"] |= np.uint32(1 << a)
        else:
            assert b > a
            bpm[a] |= np.uint32(1 << b)
    # bounding capsules: shape frame for moving shapes, world frame for static ones
    bcap = np.zeros((len(shapes), 8), dtype="<f4")
    for i, sh in enumerate(all_shapes):
        p0, p1, r = bounding_capsule(sh.desc)
        if sh.body < 0:
            R = X.quat_to_mat(sh.quat)
            p0 = R @ p0 + sh.pos; p1 = R @ p1 + sh.pos
        bcap[i, 0:3] = p0; bcap[i, 3] = r; bcap[i, 4:7] = p1
    # diagonal blocks of the mass matrix: consecutive joint dofs of one articulation
    starts = []
    last_art = None
    for i, b in enumerate(scene.bodies):
        if b.jtype == JOINT_FREE:
            continue
        if b.art != last_art:
            starts.append(b.dof); last_art = b.art
    n_block = len(starts)
    assert n_block <= 3
    starts = starts + [int(scene.header["n_jdof"])] * (4 - n_block)

    # arm-limit classifier weights (human-active ids only)
    mlp = np.zeros(0, dtype="<f4")
    mlp_dof = [-1, -1, -1, -1]
    if getattr(scene, "mlp_layers", None):
        parts = []
        for k, bb in scene.mlp_layers:
            parts += [np.asarray(k, dtype="<f4").ravel(), np.asarray(bb, dtype="<f4").ravel()]
        mlp = np.concatenate(parts)
        assert mlp.size == 8705
        for b in scene.bodies:
            if b.art == 1 and b.ref_joint in (7, 8, 9, 10):
                mlp_dof[b.ref_joint - 7] = b.dof

    # BedBathing wiping targets (bed_bathing.py:360-379)
    tgt = np.zeros((0, 4), dtype="<f4")
    n_target_upper = 0
    if getattr(scene, "targets", None) is not None:
        up, fo = scene.targets
        n_target_upper = len(up)
        tgt = np.zeros((len(up) + len(fo), 4), dtype="<f4")
        tgt[:len(up), :3] = up; tgt[len(up):, :3] = fo
        assert len(tgt) <= 160

    hdr = np.zeros(1, dtype=HEADER_DT)
    h = hdr[0]
    h["magic"] = AVG_MAGIC; h["version"] = AVG_VERSION
    vals = dict(scene.header)
    if overrides:
        vals.update(overrides)
    h["n_internal"] = 1; h["tool_body"] = -1
    h["warmstart"] = WARMSTART_FACTOR
    for k, v in vals.items():
        h[k] = v
    h["n_shape"] = n_top; h["n_cshape"] = len(cshapes); h["pshape"] = len(all_shapes) if particle is not None else -1
    h["n_mshape"] = scene.n_mshape; h["n_vert"] = len(verts); h["n_plane"] = len(planes)
    h["n_pair"] = len(pairs); h["n_frame"] = len(frames)
    h["n_block"] = n_block; h["block_start"] = starts
    h["n_mlp"] = mlp.size; h["mlp_dof"] = mlp_dof
    h["n_target"] = len(tgt); h["n_target_upper"] = n_target_upper
    off = _align(HEADER_DT.itemsize)
    sections = []
    for name, arr in (("off_body", bodies), ("off_dof", dofs), ("off_shape", shapes), ("off_vert", verts),
                      ("off_plane", planes), ("off_pair", pairs), ("off_frame", frames), ("off_bps", bps), ("off_bpm", bpm),
                      ("off_bcap", bcap), ("off_mlp", mlp), ("off_target", tgt), ("off_caabb", caabb)):
        h[name] = off
        sections.append((off, arr.tobytes()))
        off = _align(off + arr.nbytes)
    h["total_bytes"] = off
    buf = bytearray(off)
    buf[:HEADER_DT.itemsize] = hdr.tobytes()
    for o, b in sections:
        buf[o:o + len(b)] = b
    return bytes(buf)


# Warm-starting factor of the contact normal rows [UPSTREAM-BULLET, from memory]: btContactSolverInfo defaults to 0.85, the PyBullet
# physics server sets m_warmstartingFactor = 0.1 when it creates the world; the reference never touches it.
WARMSTART_FACTOR = 0.1


def read_blob(blob: bytes) -> dict:
    h = np.frombuffer(blob, dtype=HEADER_DT, count=1)[0]
    assert h["magic"] == AVG_MAGIC and h["version"] == AVG_VERSION, "bad model blob"
    out = {"header": h}
    out["bodies"] = np.frombuffer(blob, dtype=BODY_DT, count=int(h["n_body"]), offset=int(h["off_body"]))
    out["dofs"] = np.frombuffer(blob, dtype=DOF_DT, count=int(h["n_dof"]), offset=int(h["off_dof"]))
    n_all = int(h["n_shape"]) + int(h["n_cshape"]) + (1 if int(h["n_particle"]) > 0 else 0)
    out["shapes"] = np.frombuffer(blob, dtype=SHAPE_DT, count=n_all, offset=int(h["off_shape"]))
    out["caabb"] = np.frombuffer(blob, dtype="<f4", count=8 * int(h["n_cshape"]), offset=int(h["off_caabb"])).reshape(-1, 8)
    out["verts"] = np.frombuffer(blob, dtype="<f4", count=4 * int(h["n_vert"]), offset=int(h["off_vert"])).reshape(-1, 4)[:, :3]
    out["planes"] = np.frombuffer(blob, dtype="<f4", count=4 * int(h["n_plane"]), offset=int(h["off_plane"])).reshape(-1, 4)
    out["pairs"] = np.frombuffer(blob, dtype="<u4", count=int(h["n_pair"]), offset=int(h["off_pair"]))
    out["frames"] = np.frombuffer(blob, dtype=FRAME_DT, count=int(h["n_frame"]), offset=int(h["off_frame"]))
    out["bps"] = np.frombuffer(blob, dtype=BPS_DT, count=int(h["n_shape"] - h["n_mshape"]), offset=int(h["off_bps"]))
    out["bpm"] = np.frombuffer(blob, dtype="<u4", count=int(h["n_mshape"]), offset=int(h["off_bpm"]))
    out["bcap"] = np.frombuffer(blob, dtype="<f4", count=8 * int(h["n_shape"]), offset=int(h["off_bcap"])).reshape(-1, 8)
    out["mlp"] = np.frombuffer(blob, dtype="<f4", count=int(h["n_mlp"]), offset=int(h["off_mlp"]))
    out["targets"] = np.frombuffer(blob, dtype="<f4", count=4 * int(h["n_target"]), offset=int(h["off_target"])).reshape(-1, 4)[:, :3]
    return out
