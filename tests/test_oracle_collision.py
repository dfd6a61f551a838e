"""Narrowphase of the CPU oracle against closed forms / brute force (GJK closest points + margins)."""
import numpy as np
import pytest

from helpers import quat_rot, seg_seg_distance

SPHERE, CAPSULE, BOX, CYL, HULL, PLANE = 0, 1, 2, 3, 4, 5
I4 = np.array([0.0, 0, 0, 1])


@pytest.fixture(scope="module")
def shapes(oracles):
    m = oracles[0].model
    by_type = {}
    for i, s in enumerate(m["shapes"]):
        by_type.setdefault(int(s["type"]), []).append(i)
    return m, by_type


def rand_quat(rng):
    q = rng.normal(size=4)
    return q / np.linalg.norm(q)


def test_sphere_sphere(oracles, shapes):
    m, bt = shapes
    a, b = bt[SPHERE][0], bt[SPHERE][-1]
    ra, rb = float(m["shapes"][a]["radius"]), float(m["shapes"][b]["radius"])
    rng = np.random.RandomState(0)
    for _ in range(50):
        pa, pb = rng.uniform(-0.2, 0.2, 3), rng.uniform(-0.2, 0.2, 3)
        hit, out = oracles[0].shape_pair(a, np.r_[pa, I4], b, np.r_[pb, I4])
        d = np.linalg.norm(pa - pb)
        assert hit and abs(out[9] - (d - ra - rb)) < 1e-9
        n = (pa - pb) / d
        assert np.abs(out[6:9] - n).max() < 1e-7
        assert np.abs(out[0:3] - (pa - n * ra)).max() < 1e-7 and np.abs(out[3:6] - (pb + n * rb)).max() < 1e-7


def test_capsule_capsule_matches_segment_distance(oracles, shapes):
    m, bt = shapes
    a, b = bt[CAPSULE][0], bt[CAPSULE][1]
    sa, sb = m["shapes"][a], m["shapes"][b]
    rng = np.random.RandomState(1)
    for _ in range(100):
        pa, pb = rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.3, 0.3, 3)
        qa, qb = rand_quat(rng), rand_quat(rng)
        za, zb = quat_rot(qa, [0, 0, float(sa["half"][2])]), quat_rot(qb, [0, 0, float(sb["half"][2])])
        d = seg_seg_distance(pa - za, pa + za, pb - zb, pb + zb)
        if d < 1e-3:
            continue
        hit, out = oracles[0].shape_pair(a, np.r_[pa, qa], b, np.r_[pb, qb])
        assert hit and abs(out[9] - (d - float(sa["radius"]) - float(sb["radius"]))) < 1e-7
        # witness points lie on the two surfaces and are separated by dist along n
        assert np.abs((out[0:3] - out[3:6]) - out[6:9] * out[9]).max() < 1e-7


def test_sphere_box_matches_clamp_formula(oracles, shapes):
    m, bt = shapes
    s_i, b_i = bt[SPHERE][0], bt[BOX][0]
    r = float(m["shapes"][s_i]["radius"]); half = np.asarray(m["shapes"][b_i]["half"], dtype=np.float64)
    margin = float(m["shapes"][b_i]["margin"])
    rng = np.random.RandomState(2)
    for _ in range(100):
        qb = rand_quat(rng)
        local = rng.uniform(-0.15, 0.15, 3)
        core = half - margin
        cl = np.clip(local, -core, core)
        dcore = np.linalg.norm(local - cl)
        if dcore < 1e-3:
            continue
        ps = quat_rot(qb, local)
        hit, out = oracles[0].shape_pair(s_i, np.r_[ps, I4], b_i, np.r_[np.zeros(3), qb])
        # the box is Bullet's rounded box: core shrunk by the margin, margin added back as a radius
        assert hit and abs(out[9] - (dcore - margin - r)) < 1e-7


def test_hull_vs_point_matches_plane_distance(oracles, shapes):
    """A tiny sphere outside a hull, placed over the interior of a face: distance = plane distance."""
    m, bt = shapes
    s_i = bt[SPHERE][0]
    r = float(m["shapes"][s_i]["radius"])
    rng = np.random.RandomState(3)
    checked = 0
    for h_i in bt[HULL][:12]:
        hs = m["shapes"][h_i]
        verts = m["verts"][int(hs["vert_off"]):int(hs["vert_off"]) + int(hs["vert_cnt"])].astype(np.float64)
        planes = m["planes"][int(hs["plane_off"]):int(hs["plane_off"]) + int(hs["plane_cnt"])].astype(np.float64)
        centre = verts.mean(0)
        for pl in planes[:6]:
            n, d = pl[:3], pl[3]
            on = verts[np.abs(verts @ n - d) < 1e-6]
            if len(on) < 3:
                continue
            foot = on.mean(0)                       # interior point of the face
            for off in (0.003, 0.02):
                p = foot + n * (off + r + 0.001)
                # brute force: point-to-hull distance = max over planes when the foot is inside the face
                if np.max(planes[:, :3] @ (foot + n * 1e-9) - planes[:, 3]) > 1e-6:
                    continue
                hit, out = oracles[0].shape_pair(s_i, np.r_[p, I4], h_i, np.r_[np.zeros(3), I4])
                assert hit and abs(out[9] - off) < 2e-6, (h_i, out[9], off)
                assert np.abs(out[6:9] - n).max() < 1e-4
                checked += 1
    assert checked > 20


def test_contact_threshold_and_plane(oracles, shapes):
    m, bt = shapes
    s_i, p_i = bt[SPHERE][0], bt[PLANE][0]
    r = float(m["shapes"][s_i]["radius"])
    hit, out = oracles[0].shape_pair(s_i, np.r_[[0.1, 0.2, r + 0.004], I4], p_i, np.r_[np.zeros(3), I4], thr=0.005)
    assert hit and abs(out[9] - 0.004) < 1e-12 and np.allclose(out[6:9], [0, 0, 1])
    hit, _ = oracles[0].shape_pair(s_i, np.r_[[0.1, 0.2, r + 0.006], I4], p_i, np.r_[np.zeros(3), I4], thr=0.005)
    assert not hit


def test_deep_penetration_fallback_sphere_in_box(oracles, shapes):
    m, bt = shapes
    s_i, b_i = bt[SPHERE][-1], bt[BOX][0]          # tool tip sphere (r = 0.01) inside the tool handle box
    half = np.asarray(m["shapes"][b_i]["half"], dtype=np.float64)
    r = float(m["shapes"][s_i]["radius"])
    p = np.array([half[0] - 0.004, 0.0, 0.0])       # centre 4 mm inside the +x face
    hit, out = oracles[0].shape_pair(s_i, np.r_[p, I4], b_i, np.r_[np.zeros(3), I4])
    assert hit and np.allclose(out[6:9], [1, 0, 0], atol=1e-9)
    assert abs(out[9] - (-(0.004 + r))) < 1e-9


def test_scene_contacts_are_symmetric_in_sign(oracles, env_data):
    """Contacts found at reset: normal unit length, dist below the pair threshold, witness gap consistent."""
    from assistive_vr_gym_b200.compiler.reset import sample_states
    from oracle.oracle import env_to_f64
    env, var = sample_states(env_data[1], 64, np.random.RandomState(3))
    n_found = 0
    for e in range(64):
        o = oracles[int(var[e])]
        c = o.collide(env_to_f64(env[e]).copy())
        for row in c:
            n_found += 1
            assert abs(np.linalg.norm(row[8:11]) - 1) < 1e-9
            sa, sb = o.model["shapes"][int(row[0])], o.model["shapes"][int(row[1])]
            assert row[11] < min(float(sa["thr"]), float(sb["thr"]))
            assert np.abs((row[2:5] - row[5:8]) - row[8:11] * row[11]).max() < 1e-7
    assert n_found > 0


def test_gjk_coplanar_simplex_hull_vs_capsule(oracles):
    """A hull against a capsule core (a segment): the Minkowski difference contains parallelograms (two hull vertices x
    the two segment ends), so GJK's 4-point simplices are routinely coplanar.  The closest distance must not depend on
    the rounding noise of the apex-side test (regression: 1.2 cm error between the PR2 forearm hull and the human forearm).
    Checked against a brute-force minimisation over hull samples x the segment for many relative poses."""
    o = oracles[0]
    sh = o.model["shapes"]
    hulls = [i for i in range(len(sh)) if sh[i]["type"] == 4 and sh[i]["vert_cnt"] >= 8]
    caps = [i for i in range(len(sh)) if sh[i]["type"] == 1]
    assert hulls and caps
    rng = np.random.RandomState(11)
    checked = 0
    for trial in range(60):
        ia = hulls[trial % len(hulls)]; ib = caps[trial % len(caps)]
        va = o.model["verts"][int(sh[ia]["vert_off"]):int(sh[ia]["vert_off"]) + int(sh[ia]["vert_cnt"])].astype(np.float64)
        hl = float(sh[ib]["half"][2])
        size = np.linalg.norm(va.max(0) - va.min(0))
        ctr = va.mean(0)
        # capsule axis (its local z) parallel to a random hull edge direction makes coplanar simplices likely
        e = va[rng.randint(len(va))] - va[rng.randint(len(va))]
        if np.linalg.norm(e) < 1e-6:
            continue
        z = e / np.linalg.norm(e)
        x = np.cross(z, rng.normal(size=3)); x /= np.linalg.norm(x); y = np.cross(z, x)
        R = np.stack([x, y, z], axis=1)
        w = 1.0 + np.trace(R); q = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1], w]); q /= np.linalg.norm(q)
        off = rng.normal(size=3); off /= np.linalg.norm(off)
        pb = ctr + off * (0.75 * size + 0.1)
        hit, out = o.shape_pair(ia, [0, 0, 0, 0, 0, 0, 1], ib, list(pb) + list(q))
        if not hit:
            continue
        core = out[9] + float(sh[ia]["margin"]) + float(sh[ib]["margin"])
        if core <= 1e-6:
            continue
        # brute force: distance from the segment to the convex hull = min over hull surface samples; use the hull's
        # vertices, edge and face samples via random convex combinations of vertex triples
        t = np.linspace(-hl, hl, 201)
        seg = pb[None, :] + t[:, None] * z[None, :]
        lam = rng.dirichlet([0.3, 0.3, 0.3], size=4000)
        idx = rng.randint(len(va), size=(4000, 3))
        pts = np.concatenate([va, (va[idx] * lam[:, :, None]).sum(1)], axis=0)
        d = np.sqrt(((pts[:, None, :] - seg[None, :, :]) ** 2).sum(-1)).min()
        assert core <= d + 1e-9, (trial, core, d)            # GJK's distance is the global minimum: no sample may beat it
        checked += 1
    assert checked >= 30
