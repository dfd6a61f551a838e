"""Episode reset on the device (avg_reset) against its numpy mirror (compiler/reset.py sample_states_hashed), which
restates the random draws of reference ScratchItchEnv.reset."""
import numpy as np
import pytest


def test_hashed_sampler_statistics(env_data):
    """The counter-based draws have the reference's distributions (world_creation.py:66-72,141; scratch_itch.py:278)."""
    from assistive_vr_gym_b200.compiler.reset import sample_states_hashed, reset_table_bytes, RESET_TABLE_DT
    blobs, resets = env_data
    n = 20000
    env, var = sample_states_hashed(resets, n, 1234, np.ones(n, dtype=np.int64))
    assert abs(var.mean() - 0.5) < 0.02                                   # gender
    ls, st = env[:, 97], env[:, 96]
    assert abs((ls < 1).mean() - 0.25) < 0.02 and ls.min() >= 0.5        # impairment "limits": U(0.5, 1)
    assert abs((st < 1).mean() - 0.25) < 0.02 and st.min() >= 0.25       # "weakness": U(0.25, 1)
    tr = env[:, 100:110]
    assert abs((np.abs(tr).sum(1) > 0).mean() - 0.25) < 0.02 and np.abs(tr).max() <= np.deg2rad(10) + 1e-6
    assert abs((env.view(np.int32)[:, 123] == 5).mean() - 0.5) < 0.02    # limb: shoulder frame or elbow frame
    assert len(reset_table_bytes(resets[0])) == RESET_TABLE_DT.itemsize
    # a different episode index or seed gives different draws, the same one reproduces
    env2, _ = sample_states_hashed(resets, n, 1234, np.full(n, 2, dtype=np.int64))
    env3, _ = sample_states_hashed(resets, n, 1234, np.ones(n, dtype=np.int64))
    assert not np.array_equal(env, env2) and np.array_equal(env, env3)


@pytest.mark.gpu
def test_device_reset_matches_host_mirror():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from assistive_vr_gym_b200 import make
    from assistive_vr_gym_b200.compiler.reset import sample_states_hashed
    n = 4096
    env = make("ScratchItchJaco-v0", num_envs=n, device=0, seed=3, device_ik=False)     # the mirror draws start poses from the pool
    obs = env.reset_device(seed=77).cpu().numpy().copy()
    st = env.get_state()
    ref, var = sample_states_hashed(env.reset_data, n, 77, np.ones(n, dtype=np.int64))
    keep = [i for i in range(192) if i not in (162, 163, 164)]            # target_pos is filled in by the observation kernel
    assert np.array_equal(st.view(np.int32)[:, 123], ref.view(np.int32)[:, 123])
    assert np.abs(st[:, keep] - ref[:, keep]).max() < 2e-6                # sincosf / float32 rounding only
    # the observation is the one a host reset of the same records produces
    env_b = make("ScratchItchJaco-v0", num_envs=n, device=0, seed=3)
    env_b.set_state(ref, var)
    assert np.abs(env_b.obs.cpu().numpy() - obs).max() < 1e-5
    # masked reset: only the chosen environments start a new episode (episode counter 2), the others keep stepping
    a = torch.zeros((n, 7), device="cuda")
    env.step(a); env.step(a)
    before = env.get_state().copy()
    mask = torch.zeros(n, dtype=torch.bool, device="cuda"); mask[::3] = True
    env.reset_device(mask=mask)
    after = env.get_state()
    m = mask.cpu().numpy()
    assert np.array_equal(after[~m], before[~m])
    ep = np.where(m, 2, 1)
    ref2, _ = sample_states_hashed(env.reset_data, n, 77, ep)
    assert np.abs(after[m][:, keep] - ref2[m][:, keep]).max() < 2e-6
    assert (after.view(np.int32)[m, 152] == 0).all() and (after.view(np.int32)[~m, 152] == 2).all()     # iteration counters
    # stepping after a device reset works and stays finite
    for _ in range(3):
        o, r, d, info = env.step(torch.rand((n, 7), device="cuda") * 2 - 1)
    assert torch.isfinite(o).all() and torch.isfinite(r).all()
    env.close(); env_b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("env_id,ee_tol", [("ScratchItchJaco-v0", 0.03), ("ScratchItchPR2-v0", 0.03)])
def test_device_ik_start_poses(env_id, ee_tol):
    """device_ik=True: every reset environment gets its own start target (centre +- 0.05, scratch_itch.py:243,251) and an arm
    pose solved for it on the GPU.  Checked with the oracle's forward kinematics: the end-effector link reaches the drawn
    target within the reference's acceptance test (util.py:51: 0.03 in position and in quaternion distance) in >= 97 % of
    the environments (the reference, too, falls back to its closest attempt after 40 restarts), joints inside their limits,
    the tool exactly in the gripper, start poses all different, and the episode steps normally afterwards."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from assistive_vr_gym_b200 import make
    from assistive_vr_gym_b200.compiler import xform as X
    from assistive_vr_gym_b200.compiler.reset import device_ik_setup
    from oracle.oracle import Oracle, env_to_f64
    n = 2048
    env = make(env_id, num_envs=n, device=0, seed=3, device_ik=True)
    env.reset_device(seed=99)
    torch.cuda.synchronize()
    st = env.get_state(); variants = env.sim.get_variants()
    oracles = {}
    ok = 0; poses = set()
    for e in range(0, n, 8):
        v = int(variants[e])
        if v not in oracles:
            oracles[v] = (Oracle(env.blobs[v]), device_ik_setup(env.blobs[v], env.spec["task"], env.spec["robot"]))
        o, ik = oracles[v]
        rec = env_to_f64(st[e]).copy()
        target = rec[145:148]                                                 # inspection slots of avg_reset_ik_kernel (AVG_E_EBODY + 21..23)
        assert np.all(np.abs(target - ik["target_pos"]) <= 0.05 + 1e-6)
        bp = o.body_pose(rec, ik["ee_body"])
        ee_p, ee_q = X.tf_mul(bp[:3], bp[3:], ik["ee_pos"], ik["ee_quat"])
        epos = np.linalg.norm(ee_p - target)
        eq = min(np.linalg.norm(ee_q - ik["target_quat"]), np.linalg.norm(ee_q + ik["target_quat"]))
        ok += int(epos < ee_tol and eq < 0.03)
        assert abs(epos - rec[148]) < 1e-4                                 # the error the kernel reports is the real one
        d = o.model["dofs"]
        for i in range(int(o.model["header"]["n_jdof"])):
            if 0 <= d[i]["action"] < 7:
                q = rec[int(o.model["bodies"][int(d[i]["body"])]["qidx"])]
                if d[i]["lower"] <= d[i]["upper"]:
                    assert d[i]["lower"] - 1e-5 <= q <= d[i]["upper"] + 1e-5
                assert rec[64 + i] == q                                    # motor target = start pose
        wp, tb = o.frame(rec, 2), o.frame(rec, 1)
        assert np.linalg.norm(wp[:3] - tb[:3]) < 1e-5                      # tool welded where init_tool puts it
        poses.add(tuple(np.round(rec[:7], 4)))
    total = len(range(0, n, 8))
    assert ok >= 0.97 * total, (ok, total)
    assert len(poses) >= 0.95 * total
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for t in range(5):
        obs, rew, done, info = env.step(torch.rand((n, 7), device="cuda", generator=g) * 2 - 1)
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    env.close()
