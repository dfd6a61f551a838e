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
    env = make("ScratchItchJaco-v0", num_envs=n, device=0, seed=3)
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
