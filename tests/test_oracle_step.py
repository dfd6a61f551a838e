"""The oracle's env.step semantics against a plain-numpy restatement of the in-tree reference arithmetic
(take_step env.py:274-351, reward scratch_itch.py:53-82, observation scratch_itch.py:104-128, human_preferences
env.py:412-448) and against the committed oracle golden fixture."""
import os

import numpy as np
import pytest

from assistive_vr_gym_b200.compiler.reset import sample_states
from oracle.oracle import Oracle, env_to_f64
from helpers import patch_blob, quat_rot

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F_TOOL_TIP, F_TORSO, F_CHEST, F_SHOULDER, F_ELBOW, F_WRIST = 0, 3, 4, 5, 6, 7


def test_golden_fixture_reproduces(env_data, oracles):
    g = np.load(os.path.join(GOLD, "ScratchItchJaco_oracle.npz"))
    n = g["env"].shape[0]; T = g["actions"].shape[0]
    for e in range(n):
        rec = env_to_f64(g["env"][e]).copy()
        o = oracles[int(g["variant"][e])]
        assert np.abs(o.reset_obs(rec) - g["obs0"][e]).max() < 1e-12
        for t in range(T):
            obs, rew, info, c = o.step(rec, g["actions"][t, e])
            assert np.abs(obs - g["obs"][t, e]).max() < 1e-9
            assert abs(rew - g["reward"][t, e]) < 1e-9
            assert len(c) == g["ncontacts"][t, e]
            assert np.abs(rec - g["states"][t, e]).max() < 1e-9


def test_motor_targets_follow_take_step(env_data, oracles):
    """env.py:275-326: clip to [-1,1], x0.05, five limit-masked accumulations (mask sticky)."""
    blobs, resets = env_data
    env, var = sample_states(resets, 6, np.random.RandomState(11), genders=np.zeros(6, dtype=np.int32))
    o = oracles[0]
    dofs = o.model["dofs"]
    rd = resets[0]
    for e in range(6):
        rec = env_to_f64(env[e]).copy()
        rec[rd["arm_qidx"][1]] = float(dofs[1]["rep_upper"]) - 0.06          # joint 2 close to its upper limit
        action = np.array([2.0, 1.0, -3.0, 0.3, -0.2, 0.0, 0.9], dtype=np.float32)
        a = np.clip(action, -1, 1) * np.float32(0.05)
        pos = rec[rd["arm_qidx"]].copy()
        a = a.astype(np.float64)
        lo = np.array([dofs[i]["rep_lower"] for i in rd["arm_dof"]], dtype=np.float64)
        hi = np.array([dofs[i]["rep_upper"] for i in rd["arm_dof"]], dtype=np.float64)
        for _ in range(5):
            a[pos + a < lo] = 0
            a[pos + a > hi] = 0
            pos += a
        o.step(rec, action)
        assert np.abs(rec[64 + rd["arm_dof"]] - pos).max() < 1e-12
        assert abs(pos[1] - (float(dofs[1]["rep_upper"]) - 0.06 + 0.05)) < 1e-6   # one increment fits, the second is masked


def test_tremor_targets_alternate(env_data, oracles):
    """env.py:330-333: with the tremor impairment the human targets are target +- tremor, sign flipping each step."""
    blobs, resets = env_data
    env, var = sample_states(resets, 200, np.random.RandomState(12), genders=np.zeros(200, dtype=np.int32))
    e = int(np.nonzero(env[:, 99] > 0)[0][0])
    rec = env_to_f64(env[e]).copy()
    rd = resets[0]
    slots = rd["hum_joint"] - 4
    base = rec[110 + slots].copy(); trem = rec[100 + slots].copy()
    o = oracles[0]
    for it in range(4):
        o.step(rec, np.zeros(7, dtype=np.float32))
        sgn = 1.0 if it % 2 == 0 else -1.0
        assert np.abs(rec[64 + rd["hum_dof"]] - (base + trem * sgn)).max() < 1e-12
        assert rec[98] == pytest.approx(0.05)                       # human_gains passed by scratch_itch.py:45


def test_reward_and_obs_follow_reference_formulas(env_data, oracles):
    """Recompute reward terms and the observation from oracle state with the reference's formulas."""
    blobs, resets = env_data
    g = np.load(os.path.join(GOLD, "ScratchItchJaco_oracle.npz"))
    w = dict(distance=1.0, action=0.01, tool_force=0.01, scratch=2.0, C_v=0.25, C_f=0.01, C_hf=0.05)   # config.ini
    for e in range(g["env"].shape[0]):
        rec = env_to_f64(g["env"][e]).copy()
        o = oracles[int(g["variant"][e])]
        rd = resets[int(g["variant"][e])]
        prev_success = 0.0
        for t in range(6):
            action = g["actions"][t, e]
            obs, rew, info, c = o.step(rec, action)
            tool = o.frame(rec, F_TOOL_TIP); torso = o.frame(rec, F_TORSO)
            sh, el, wr = o.frame(rec, F_SHOULDER), o.frame(rec, F_ELBOW), o.frame(rec, F_WRIST)
            limb = o.frame(rec, int(rec[123]))
            target = limb[:3] + quat_rot(limb[3:], rec[120:123])
            expect = np.concatenate([tool[:3] - torso[:3], tool[3:], tool[:3] - target, target - torso[:3], rec[rd["arm_qidx"]],
                                     sh[:3] - torso[:3], el[:3] - torso[:3], wr[:3] - torso[:3], [info[2]]])
            assert np.abs(obs - expect).max() < 1e-12
            # reward, scratch_itch.py:62-72 + env.py:412-448
            tqd = rec[32 + 17:32 + 23]
            tb = o.body_pose(rec, 17)
            ee_vel = np.linalg.norm(tqd[:3] + np.cross(tqd[3:], tool[:3] - tb[:3]))
            total, tf_target = info[0], info[3]
            pref = w["C_v"] * -ee_vel + w["C_f"] * -(total - tf_target) + w["C_hf"] * (0 if tf_target < 10 else -tf_target)
            r = (w["distance"] * -np.linalg.norm(target - tool[:3]) + w["action"] * -np.sum(np.square(action.astype(np.float64)))
                 + w["tool_force"] * tf_target + w["scratch"] * info[6] + pref)
            assert abs(r - rew) < 1e-7          # config weights are stored as float32 in the blob
            assert info[5] == pytest.approx(-np.sum(np.square(action.astype(np.float64))))   # raw, unclipped action
            assert rec[153] >= prev_success
            prev_success = rec[153]


def test_forces_come_from_tool_and_robot_contacts_only(env_data, oracles):
    """scratch_itch.py:84-102: tool_force sums every tool contact, total_force_on_human sums tool-human and robot-human."""
    blobs, resets = env_data
    env, var = sample_states(resets, 64, np.random.RandomState(3))
    seen = 0
    for e in range(64):
        o = oracles[int(var[e])]
        rec = env_to_f64(env[e]).copy()
        obs, rew, info, c = o.step(rec, np.zeros(7, dtype=np.float32))
        sh = o.model["shapes"]
        tool = total = 0.0
        for row in c:
            ra, rb = int(sh[int(row[0])]["ref_body"]), int(sh[int(row[1])]["ref_body"])
            f = row[12]
            assert f >= 0
            if 2 in (ra, rb):
                tool += f
            if {ra, rb} == {2, 1} or {ra, rb} == {0, 1}:
                total += f
                seen += 1
        assert info[2] == pytest.approx(tool) and info[0] == pytest.approx(total)
    assert seen > 0


def test_warm_start_cache_round_trip(env_data, oracles):
    """AVG_E_WCACHE (round 2): after a step with contacts the record holds (pair key, normal impulse) of the first contact points;
    the next internal step starts those rows from warmstart x impulse (btMultiBodyConstraintSolver::setupMultiBodyContactConstraint).
    Checked on the oracle: the keys are the contact pairs, the impulses the reported forces x dt, and switching the factor off
    (header.warmstart = 0) changes the next step only through those rows."""
    from assistive_vr_gym_b200.compiler.reset import sample_states
    from oracle.oracle import Oracle, env_to_f64
    from helpers import patch_blob
    blobs, resets = env_data
    env0, variant = sample_states(resets, 200, np.random.RandomState(11))
    rng = np.random.RandomState(2)
    found = 0
    for e in range(200):
        o = oracles[int(variant[e])]
        rec = env_to_f64(env0[e]).copy()
        drift = rng.uniform(-0.8, 0.8, 7)
        for t in range(25):
            _, _, _, oc = o.step(rec, np.clip(drift + rng.uniform(-0.4, 0.4, 7), -1, 1))
            if len(oc):
                break
        if not len(oc):
            continue
        found += 1
        dt = float(o.model["header"]["dt"])
        keys = [int(rec[176 + 2 * w]) for w in range(8)]; imps = [rec[177 + 2 * w] for w in range(8)]
        for w, c in enumerate(oc[:8]):
            assert keys[w] == (int(c[0]) | (int(c[1]) << 16)) and abs(imps[w] - c[12] * dt) < 1e-12
        assert all(k == 0 for k in keys[len(oc):])
        # the same next step with and without the cache: identical unless a cached pair is in contact again
        a = rng.uniform(-1, 1, 7)
        warm1 = Oracle(patch_blob(blobs[int(variant[e])], header={"substeps": 1}))                      # one internal step per env-step
        cold1 = Oracle(patch_blob(blobs[int(variant[e])], header={"substeps": 1, "warmstart": 0.0}))
        r1 = rec.copy(); r2 = rec.copy(); r2[176:192] = 0.0; r3 = rec.copy()
        _, _, _, c1 = warm1.step(r1, a); warm1.step(r2, a); cold1.step(r3, a)
        assert np.array_equal(r2[:64], r3[:64])                      # empty cache == factor 0
        if any((int(c[0]) | (int(c[1]) << 16)) in keys for c in c1):
            assert not np.array_equal(r1[32:64], r2[32:64])          # a cached pair in contact again: its row started from 0.1 x impulse
        if found >= 6:
            break
    assert found >= 3
