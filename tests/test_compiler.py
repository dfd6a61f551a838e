"""Model-compiler checks against facts hard-coded in the reference (needs the reference asset tree)."""
import os

import numpy as np
import pytest

from conftest import ASSETS

pytestmark = pytest.mark.assets


@pytest.fixture(scope="module")
def scene():
    from assistive_vr_gym_b200.compiler.scene import build_scratch_itch
    return build_scratch_itch(ASSETS, "jaco", "male")


def test_jaco_link_numbering_matches_reference_indices():
    """Jaco arm joints 1..7, end effector link 8, finger joints 9/11/13 (world_creation.py:283,320,332)."""
    from assistive_vr_gym_b200.compiler.urdf import parse_urdf
    u = parse_urdf(os.path.join(ASSETS, "jaco", "j2s7s300_gym.urdf"))
    names = [j.child for j in u.order]
    assert names[1:8] == [f"j2s7s300_link_{i}" for i in range(1, 8)]
    assert names[8] == "j2s7s300_end_effector"
    assert [names[i] for i in (9, 11, 13)] == [f"j2s7s300_link_finger_{i}" for i in (1, 2, 3)]
    assert len(names) == 15                          # the wheelchair_link block is commented out in the URDF
    movable = [i for i, j in enumerate(u.order) if j.jtype != "fixed"]
    assert movable == [1, 2, 3, 4, 5, 6, 7, 9, 11, 13]


def test_dae_meshes_match_their_stl_twins():
    from assistive_vr_gym_b200.compiler.meshes import load_dae, load_stl
    d = os.path.join(ASSETS, "jaco", "meshes")
    for f in ("base", "shoulder", "arm_half_1", "arm_half_2", "forearm", "wrist_spherical_1", "wrist_spherical_2", "hand_3finger"):
        a, b = load_dae(os.path.join(d, f + ".dae")), load_stl(os.path.join(d, f + ".STL"))
        assert np.abs(a.min(0) - b.min(0)).max() < 2e-4 and np.abs(a.max(0) - b.max(0)).max() < 2e-4, f


def test_human_link_legend():
    """human_creation.py:27-45: which link indices carry the body parts after PyBullet's depth-first renumbering."""
    from assistive_vr_gym_b200.compiler.human import create_human
    from assistive_vr_gym_b200.compiler.mbody import SHAPE_CAPSULE, SHAPE_SPHERE, SHAPE_HULL
    h = create_human(ASSETS, "male", 0.6)
    with_shape = [l.ref_index for l in h.links if l.shapes]
    assert with_shape == [2, 3, 6, 9, 11, 13, 16, 19, 21, 23, 24, 27, 30, 31, 34, 37, 38, 41]
    assert h.links[13].shapes[0].kind == SHAPE_SPHERE and h.links[9].shapes[0].kind == SHAPE_CAPSULE
    assert all(s.kind == SHAPE_HULL for s in h.links[27].shapes) and len(h.links[27].shapes) == 8
    rev = [l.ref_index for l in h.links if l.jtype == "revolute"]
    assert rev == list(range(7, 14)) + list(range(17, 24)) + list(range(24, 42))      # joints 0-6, 14-16 are fixed
    assert h.links[9].mass == pytest.approx(78.4 * 0.033) and h.links[11].mass == pytest.approx(78.4 * 0.019)
    # right arm limits, human_creation.py:226-228
    assert np.rad2deg(h.links[7].lower) == pytest.approx(5) and np.rad2deg(h.links[10].lower) == pytest.approx(-128)
    f = create_human(ASSETS, "female", 0.54)
    assert len(f.links[27].shapes) == 9 and f.links[9].shapes[0].radius == pytest.approx(0.0355)


def test_scene_structure(scene):
    assert len(scene.bodies) == 18 and len(scene.dofs) == 23 and scene.n_mshape == 19
    art = [b.art for b in scene.bodies]
    assert art == [0] * 10 + [1] * 7 + [2]
    # frozen human: only joints 7..13 move (world_creation.py:157-161 with controllable 4..13)
    assert [b.ref_joint for b in scene.bodies if b.art == 1] == list(range(7, 14))
    tool = scene.bodies[-1]
    assert tool.mass == pytest.approx(0.11)            # 0.05 + 0.05 + 0.01, tool_scratch.urdf
    # tool/robot filter: robot links 7..14 never collide with the tool (world_creation.py:352-354)
    for a, b in scene.pairs:
        sa, sb = scene.shapes[a], scene.shapes[b]
        refs = {sa.ref_body: sa.ref_link, sb.ref_body: sb.ref_link}
        if set(refs) == {0, 2}:
            assert refs[0] not in range(7, 15)
        if sa.mb_index == sb.mb_index == 1:            # human self collision matrix
            from assistive_vr_gym_b200.compiler.human import human_self_collision_enabled
            assert human_self_collision_enabled(sa.ref_link, sb.ref_link)
    assert scene.info["n_pairs"] == len(scene.pairs) > 2000


def test_committed_data_matches_a_fresh_compile(scene):
    from assistive_vr_gym_b200.compiler.blob import scene_to_blob
    from assistive_vr_gym_b200.envs import load_env_data
    blobs, _ = load_env_data("ScratchItchJaco.npz")
    assert scene_to_blob(scene) == blobs[0]


def test_oracle_fk_matches_numpy_fk(scene):
    """Independent check of reduce_bodies + the oracle's FK: COM frames of every moving reference link."""
    from assistive_vr_gym_b200.compiler.blob import scene_to_blob
    from oracle.oracle import Oracle
    from helpers import quat_rot
    o = Oracle(scene_to_blob(scene))
    rng = np.random.RandomState(0)
    robot, human = scene.multibodies[0], scene.multibodies[1]
    for _ in range(5):
        rec = np.zeros(192)
        qr, qh = {}, dict(scene.q_human_reset)
        for b in scene.bodies:
            if b.jtype == 2:
                rec[b.qidx:b.qidx + 7] = [0, 0, 1, 0, 0, 0, 1]
                continue
            v = rng.uniform(-1.5, 1.5)
            rec[b.qidx] = v
            (qr if b.art == 0 else qh)[b.ref_joint] = v
        for mb_i, mb, q in ((0, robot, qr), (1, human, qh)):
            com = mb.com_frames(q)
            for link in range(len(mb.links)):
                at = scene.attach[mb_i][link]
                if at.body < 0:
                    continue
                bp = o.body_pose(rec, at.body)
                l = mb.links[link]
                p = bp[:3] + quat_rot(bp[3:], at.pos + quat_rot(at.quat, l.inertial_pos))
                assert np.abs(p - com[link][0]).max() < 5e-7, (mb.name, link)      # blob transforms are float32


@pytest.mark.assets
@pytest.mark.parametrize("robot_type", ["sawyer", "baxter"])
def test_sawyer_baxter_loaders_match_reference_constants(robot_type):
    """`init_sawyer` / `init_baxter` (world_creation.py:219-272): PyBullet numbering re-derived from the URDFs reproduces the
    hard-coded arm / gripper / tool-link indices (:235,254-255,313-318,332-334); the multibody reduces to <= 32 one-lane dofs
    with a 6-dof tool (no reference environment uses these robots, SURVEY.md F4 -- they are compiled for the build-defined
    Feeding / Drinking ids)."""
    from assistive_vr_gym_b200.compiler.scene import load_robot
    from assistive_vr_gym_b200.compiler.mbody import reduce_bodies
    robot, rs = load_robot(ASSETS, robot_type)
    names = {l.ref_index: l.name for l in robot.links}
    movable = [l.ref_index for l in robot.links if l.jtype in ("revolute", "prismatic")]
    if robot_type == "sawyer":
        assert len(robot.links) == 24 and len(movable) == 10          # 8 revolute + 2 prismatic
        assert rs["arm"] == [3, 8, 9, 10, 11, 13, 16] and [names[j] for j in rs["arm"]] == [f"right_l{k}" for k in range(7)]
        assert names[18] == "right_gripper_base" and [names[j] for j in rs["fingers"]] == ["r_gripper_l_finger", "r_gripper_r_finger"]
        assert robot.self_filter(5, 12) and robot.self_filter(1, 7) and not robot.self_filter(1, 12)
    else:
        assert len(robot.links) == 56 and len(movable) == 19          # 15 revolute + 4 prismatic
        assert rs["arm"] == [12, 13, 14, 15, 16, 18, 19] and names[12] == "right_upper_shoulder" and names[19] == "right_wrist"
        assert names[25] == "right_gripper_base" and names[47] == "left_gripper_base" and names[34] == "left_upper_shoulder"
        assert [names[j] for j in rs["fingers"]] == ["r_gripper_l_finger", "r_gripper_r_finger"]
    bodies, attach = reduce_bodies(robot, 0, rs["q_preset"], rs["frozen"], 0)
    assert [b.ref_joint for b in bodies] == sorted(rs["arm"] + rs["fingers"])
    assert len(bodies) + 6 + 4 <= 32                                  # + tool + head joints of Feeding / Drinking
    assert all(b.mass > 0 and np.all(np.asarray(b.inertia) > 0) for b in bodies)
    for b in bodies:
        assert b.limit_enforced and b.lower < b.upper
