"""On-device policy inference (avg_policy_act) against a numpy restatement of the reference's rollout loop
(enjoy_vr.py:105-117: VecNormalize'd observations -> a2c_ppo_acktr MLP actor -> deterministic mean action)."""
import struct

import numpy as np
import pytest


def numpy_act(arrs, obs, n_act, clip=10.0, eps=1e-8):
    x = np.clip((obs[:, :len(arrs["ob_mean"])].astype(np.float64) - arrs["ob_mean"]) / np.sqrt(arrs["ob_var"].astype(np.float64) + eps), -clip, clip)
    h = np.tanh(x @ arrs["W1"].astype(np.float64) + arrs["b1"])
    h = np.tanh(h @ arrs["W2"].astype(np.float64) + arrs["b2"])
    a = h @ arrs["W3"].astype(np.float64) + arrs["b3"]
    out = np.zeros((obs.shape[0], n_act)); out[:, :a.shape[1]] = a
    return out


def test_policy_blob_layout():
    from assistive_vr_gym_b200.policy import synthetic_policy, POLICY_MAGIC
    blob, arrs = synthetic_policy(30, 7, seed=1)
    magic, n_in, n_out, clip, eps = struct.unpack_from("<Iiiff", blob, 0)
    assert (magic, n_in, n_out) == (POLICY_MAGIC, 30, 7) and clip == 10.0
    assert len(blob) == 32 + 4 * (2 * 30 + 30 * 64 + 64 + 4096 + 64 + 64 * 7 + 7)
    assert np.allclose(arrs["W2"].T @ arrs["W2"], 2.0 * np.eye(64), atol=1e-4)           # orthogonal, gain sqrt(2)


def test_state_dict_import_transposes():
    from assistive_vr_gym_b200.policy import from_state_dict, pack_policy
    rng = np.random.RandomState(0)
    sd = {"base.actor.0.weight": rng.normal(size=(64, 30)), "base.actor.0.bias": rng.normal(size=64),
          "base.actor.2.weight": rng.normal(size=(64, 64)), "base.actor.2.bias": rng.normal(size=64),
          "dist.fc_mean.weight": rng.normal(size=(7, 64)), "dist.fc_mean.bias": rng.normal(size=7)}
    mean, var = rng.normal(size=30), rng.uniform(0.5, 2, size=30)
    a = from_state_dict(sd, mean, var)
    b = pack_policy(mean, var, sd["base.actor.0.weight"].T, sd["base.actor.0.bias"], sd["base.actor.2.weight"].T, sd["base.actor.2.bias"],
                    sd["dist.fc_mean.weight"].T, sd["dist.fc_mean.bias"])
    assert a == b


@pytest.mark.gpu
@pytest.mark.parametrize("env_id,n_in,n_out", [("ScratchItchJaco-v0", 30, 7), ("ScratchItchJacoHuman-v0", 30, 7)])
def test_gpu_policy_matches_numpy(env_id, n_in, n_out):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from assistive_vr_gym_b200 import make
    from assistive_vr_gym_b200.policy import pack_policy, synthetic_policy
    n = 2048
    env = make(env_id, num_envs=n, device=0, seed=5)
    _, arrs = synthetic_policy(n_in, n_out, seed=3)
    rng = np.random.RandomState(4)
    arrs["ob_mean"] = rng.normal(scale=0.3, size=n_in).astype(np.float32); arrs["ob_var"] = rng.uniform(0.01, 2.0, size=n_in).astype(np.float32)
    arrs["b1"] = rng.normal(scale=0.1, size=64).astype(np.float32); arrs["b3"] = rng.normal(scale=0.1, size=n_out).astype(np.float32)
    env.set_policy(pack_policy(**arrs))
    obs = env.reset().cpu().numpy().copy()
    act = env.act().cpu().numpy()
    ref = numpy_act(arrs, obs, env.sim.n_actions)
    assert act.shape == ref.shape and np.abs(act - ref).max() < 2e-5
    if env.sim.n_actions > n_out:
        assert not act[:, n_out:].any()                                                  # human half zero, enjoy_vr.py:112-113
    # a short closed-loop rollout stays finite and equals act -> step done by hand
    env2 = make(env_id, num_envs=n, device=0, seed=5)
    env2.set_policy(pack_policy(**arrs)); env2.set_state(env.get_state(), env.variants)
    env.rollout(4)
    for _ in range(4):
        env2.step(torch.as_tensor(numpy_act(arrs, env2.obs.cpu().numpy(), env2.sim.n_actions).astype(np.float32), device="cuda"))
    assert torch.isfinite(env.obs).all()
    assert np.abs(env.get_state()[:, :64] - env2.get_state()[:, :64]).max() < 5e-3       # float32 device actions vs float64 host actions
    env.close(); env2.close()
