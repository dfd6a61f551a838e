"""Episode logs (setup.pkl / actions.pkl, reference scratch_itch.py:47-51,269-272) and replay / re-scoring
(replay_vr_savemeta.py:17-59): file formats on the CPU with a stub environment, bit-exact replay on the GPU."""
import os
import pickle

import numpy as np
import pytest


class _StubSim:
    def get_variants(self):
        return np.array([0, 1, 1], dtype=np.int32)


class _StubEnv:
    env_id = "ScratchItchJaco-v0"; spec = {"robot": "jaco"}; num_envs = 3; blobs = [b"", b""]; variants = None; sim = _StubSim()

    def get_state(self):
        return np.arange(3 * 192, dtype=np.float32).reshape(3, 192)


def test_episode_log_formats(tmp_path):
    from assistive_vr_gym_b200.replay import EpisodeRecorder, load_episode
    d = str(tmp_path / "participant_0" / "scratch_itch_jaco_run")
    rec = EpisodeRecorder(_StubEnv(), d, env_index=2)
    rec.start()
    acts = [np.random.RandomState(k).uniform(-1, 1, (3, 7)).astype(np.float32) for k in range(5)]
    for a in acts:
        rec.record(a)
    rec.close()
    with open(os.path.join(d, "setup.pkl"), "rb") as f:
        setup = pickle.load(f)
    assert setup == ["jaco", "female", 0.54]                       # [robot_type, gender, hipbone_to_mouth_height], scratch_itch.py:272
    with open(os.path.join(d, "actions.pkl"), "rb") as f:
        al = pickle.load(f)
    assert isinstance(al, list) and len(al) == 5 and all(a.shape == (7,) for a in al)     # action_list, scratch_itch.py:46,51
    assert np.array_equal(al[3], acts[3][2])
    env_id, record, variant, setup2, al2 = load_episode(d)
    assert env_id == "ScratchItchJaco-v0" and variant == 1 and setup2 == setup and np.array_equal(record, _StubEnv().get_state()[2])


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", ["ScratchItchJaco-v0", "BedBathingPR2-v0"])
def test_gpu_replay_reproduces_recorded_episode(tmp_path, env_id):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from assistive_vr_gym_b200 import make
    from assistive_vr_gym_b200.replay import EpisodeRecorder, rescore_directory
    n, k, T = 16, 5, 30
    env = make(env_id, num_envs=n, device=0, seed=11)
    env.reset_device(seed=5)                                         # variants live on the device only: avg_get_variants
    d = str(tmp_path / "participant_3" / ("bed_bathing_pr2" if "Bed" in env_id else "scratch_itch_jaco"))
    rec = EpisodeRecorder(env, d, env_index=k)
    rec.start()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    rewards, forces, obs_last = [], [], None
    for t in range(T):
        a = torch.rand((n, 7), device="cuda", generator=g) * 2 - 1
        rec.record(a)
        obs, rew, done, info = env.step(a)
        rewards.append(float(rew[k])); forces.append(float(info["total_force_on_human"][k])); obs_last = obs[k].cpu().numpy()
    rec.close()
    env.close()
    names, obs_all, rew_all, act_all, force_all, succ_all = rescore_directory(str(tmp_path), out=str(tmp_path / "observations_vr.pkl"))
    assert names == [d] and len(rew_all[0]) == T
    assert rew_all[0] == rewards and force_all[0] == forces          # deterministic kernels: bit-exact re-simulation
    assert np.array_equal(obs_all[0][-1].astype(np.float32), obs_last)
    with open(str(tmp_path / "observations_vr.pkl"), "rb") as f:
        assert len(pickle.load(f)) == 6                              # replay_vr_savemeta.py:59
