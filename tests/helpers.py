"""Shared helpers for the test-suite (pure numpy)."""
from __future__ import annotations

import numpy as np

from assistive_vr_gym_b200.compiler.blob import HEADER_DT, SHAPE_DT, DOF_DT, BODY_DT


def patch_blob(blob: bytes, header: dict | None = None, friction: float | None = None, dof_flags_clear: int = 0,
               zero_gravity: bool = False) -> bytes:
    """Return a copy of a ModelBlob with some fields overridden (test scenarios)."""
    b = bytearray(blob)
    h = np.frombuffer(b, dtype=HEADER_DT, count=1)
    for k, v in (header or {}).items():
        h[k] = v
    if friction is not None:
        sh = np.frombuffer(b, dtype=SHAPE_DT, count=int(h["n_shape"][0]), offset=int(h["off_shape"][0]))
        sh["friction"] = friction
    if dof_flags_clear:
        d = np.frombuffer(b, dtype=DOF_DT, count=int(h["n_dof"][0]), offset=int(h["off_dof"][0]))
        d["flags"] &= ~np.uint32(dof_flags_clear)
    if zero_gravity:
        bd = np.frombuffer(b, dtype=BODY_DT, count=int(h["n_body"][0]), offset=int(h["off_body"][0]))
        bd["gravity"] = 0
    return bytes(b)


def quat_rot(q, v):
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return R @ np.asarray(v, dtype=np.float64)


def seg_seg_distance(p1, q1, p2, q2):
    """Closest distance between segments [p1,q1] and [p2,q2] by dense sampling + local refinement (test oracle)."""
    d1, d2 = q1 - p1, q2 - p2
    r = p1 - p2
    a, e, f = d1 @ d1, d2 @ d2, d2 @ r
    if a <= 1e-12 and e <= 1e-12:
        return np.linalg.norm(r)
    if a <= 1e-12:
        s, t = 0.0, np.clip(f / e, 0, 1)
    else:
        c = d1 @ r
        if e <= 1e-12:
            t, s = 0.0, np.clip(-c / a, 0, 1)
        else:
            b = d1 @ d2
            den = a * e - b * b
            s = np.clip((b * f - c * e) / den, 0, 1) if den > 1e-12 else 0.0
            t = (b * s + f) / e
            if t < 0:
                t, s = 0.0, np.clip(-c / a, 0, 1)
            elif t > 1:
                t, s = 1.0, np.clip((b - c) / a, 0, 1)
    return np.linalg.norm((p1 + d1 * s) - (p2 + d2 * t))
