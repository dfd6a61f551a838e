"""Invariants of the host-side reset sampler (restating scratch_itch.py:155-162,230-256,275-287)."""
import numpy as np

from assistive_vr_gym_b200.compiler.reset import sample_states, E_Q, E_MTARGET, E_STRENGTH, E_LIMIT_SCALE, E_HUMAN_KP, \
    E_TREMOR_ON, E_TREMOR, E_TARGET_ON_ARM, E_LIMB_FRAME, F_SHOULDER, F_ELBOW


def test_sampled_states(env_data):
    blobs, resets = env_data
    env, var = sample_states(resets, 4000, np.random.RandomState(0))
    iv = env.view(np.int32)
    assert set(np.unique(var)) == {0, 1}
    assert 0.4 < var.mean() < 0.6                                   # gender choice, scratch_itch.py:156
    for v in (0, 1):
        rd = resets[v]; sel = var == v
        q = env[sel]
        arm = q[:, E_Q + rd["arm_qidx"]]
        assert np.all(arm >= -2 * np.pi - 1e-5) and np.all(arm <= 2 * np.pi + 1e-5)
        assert np.allclose(q[:, E_Q + rd["fin_qidx"]], 1.0)         # gripper open, scratch_itch.py:254
        tq = int(rd["tool_qidx"])
        assert np.allclose(np.linalg.norm(q[:, E_Q + tq + 3:E_Q + tq + 7], axis=1), 1.0, atol=1e-5)
        ls = q[:, E_LIMIT_SCALE][:, None]
        hq = q[:, E_Q + rd["hum_qidx"]]
        assert np.all(hq >= rd["hum_lower"][None] * ls - 1e-6) and np.all(hq <= rd["hum_upper"][None] * ls + 1e-6)
        assert np.allclose(q[:, E_MTARGET + rd["hum_dof"]], hq)
        # target lies on the surface of the chosen limb capsule (util.py:112-132)
        limb = np.where(iv[sel, E_LIMB_FRAME] == F_SHOULDER, 0, 1)
        assert set(np.unique(iv[sel, E_LIMB_FRAME])) == {F_SHOULDER, F_ELBOW}
        t = q[:, E_TARGET_ON_ARM:E_TARGET_ON_ARM + 3]
        rad = rd["limb_dims"][limb, 1]; length = rd["limb_dims"][limb, 0]
        assert np.allclose(np.hypot(t[:, 0], t[:, 1]), rad, atol=1e-6)
        assert np.all(-t[:, 2] >= rad - 1e-6) and np.all(-t[:, 2] <= length + 1e-6)
    # impairments (world_creation.py:66-72)
    assert np.all((env[:, E_LIMIT_SCALE] >= 0.5) & (env[:, E_LIMIT_SCALE] <= 1.0))
    assert np.all((env[:, E_STRENGTH] >= 0.25) & (env[:, E_STRENGTH] <= 1.0))
    trem = env[:, E_TREMOR_ON] > 0
    assert 0.15 < trem.mean() < 0.35
    assert np.all(np.abs(env[:, E_TREMOR:E_TREMOR + 10]) <= np.deg2rad(10) + 1e-6)
    assert np.all(env[~trem, E_TREMOR:E_TREMOR + 10] == 0)
    assert np.allclose(env[trem, E_HUMAN_KP], 0.05) and np.allclose(env[~trem, E_HUMAN_KP], 0.01)


def test_sampler_is_deterministic(env_data):
    a, va = sample_states(env_data[1], 100, np.random.RandomState(42))
    b, vb = sample_states(env_data[1], 100, np.random.RandomState(42))
    assert np.array_equal(a, b) and np.array_equal(va, vb)
