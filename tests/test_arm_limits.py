"""enforce_realistic_human_joint_limits (reference env.py:353-387): the Keras arm-limit classifier
(`realistic_arm_limits_model.h5`, Dense 4-64-64-64-1, tanh / sigmoid) behind the human-active ids.

CPU: the HDF5 subset reader against the file's own structure, the oracle's MLP against a plain-numpy restatement of
`Sequential.predict_classes` on the weights stored in the ModelBlob.  GPU: the CUDA classifier against the oracle
(logits within 2e-3, classes identical away from the decision boundary), and a sub-step in which the teleport back
to the last valid pose fires."""
import os

import numpy as np
import pytest

from conftest import ASSETS

H5 = os.path.join(ASSETS, "realistic_arm_limits_model.h5")


@pytest.fixture(scope="module")
def human_data():
    from assistive_vr_gym_b200.envs import load_env_data
    return load_env_data("ScratchItchJacoHuman.npz")


def numpy_logit(mlp, q4):
    """env.py:360-364 on raw joint angles (tz, tx, ty, qe); float64 matmuls on the float32 weights."""
    mlp = np.asarray(mlp, dtype=np.float64)
    W1, b1 = mlp[:256].reshape(4, 64), mlp[256:320]
    W2, b2 = mlp[320:4416].reshape(64, 64), mlp[4416:4480]
    W3, b3 = mlp[4480:8576].reshape(64, 64), mlp[8576:8640]
    W4, b4 = mlp[8640:8704], mlp[8704]
    tz, tx, ty, qe = q4
    x = np.array([(-tz + 2 * np.pi) % (2 * np.pi), (tx + 2 * np.pi) % (2 * np.pi), -ty, (-qe + 2 * np.pi) % (2 * np.pi)])
    h = np.tanh(x @ W1 + b1); h = np.tanh(h @ W2 + b2); h = np.tanh(h @ W3 + b3)
    return float(h @ W4 + b4)


@pytest.mark.assets
def test_h5_reader_recovers_the_dense_stack(human_data):
    from assistive_vr_gym_b200.compiler.h5lite import H5Lite, load_keras_dense_stack
    layers = load_keras_dense_stack(H5)
    assert [k.shape for k, _ in layers] == [(4, 64), (64, 64), (64, 64), (64, 1)]
    assert [b.shape for _, b in layers] == [(64,), (64,), (64,), (1,)]
    h = H5Lite(H5)
    assert sorted(h.children(h.get("model_weights"))) == ["dense_1", "dense_2", "dense_3", "dense_4"]
    # Glorot-uniform initialised, then trained: finite, O(1) weights; and the blob carries exactly these numbers
    flat = np.concatenate([np.concatenate([k.ravel(), b.ravel()]) for k, b in layers])
    assert np.isfinite(flat).all() and 0.05 < np.abs(flat).mean() < 2.0
    from assistive_vr_gym_b200.compiler.blob import read_blob
    for blob in human_data[0]:
        assert np.array_equal(read_blob(blob)["mlp"], flat)


def test_robot_only_blob_has_no_classifier(env_data):
    from assistive_vr_gym_b200.compiler.blob import read_blob
    for blob in env_data[0]:
        assert int(read_blob(blob)["header"]["n_mlp"]) == 0


def test_oracle_mlp_matches_numpy(human_data):
    from oracle.oracle import Oracle
    o = Oracle(human_data[0][0])
    md = [int(d) for d in o.model["header"]["mlp_dof"]]
    lo = np.array([o.model["dofs"][d]["lower"] for d in md]); hi = np.array([o.model["dofs"][d]["upper"] for d in md])
    rng = np.random.RandomState(0)
    qs = rng.uniform(lo - 0.5, hi + 0.5, (500, 4))
    lg = np.array([o.arm_limit_logit(q) for q in qs]); ref = np.array([numpy_logit(o.model["mlp"], q) for q in qs])
    assert np.abs(lg - ref).max() < 1e-9
    frac = (lg > 0).mean()
    assert 0.2 < frac < 0.8                      # the classifier separates the limit box; neither class is degenerate


def test_reset_poses_are_valid(human_data):
    """scratch_itch.py:230-235 starts the arm in a comfortable pose: the classifier must accept it (otherwise
    right_arm_previous_valid_pose would stay None and the limit would never engage)."""
    from assistive_vr_gym_b200.compiler.reset import sample_states
    from oracle.oracle import Oracle
    blobs, resets = human_data
    env, var = sample_states(resets, 32, np.random.RandomState(4))
    for e in range(32):
        o = Oracle(blobs[int(var[e])])
        md = [int(d) for d in o.model["header"]["mlp_dof"]]
        q4 = [env[e, int(o.model["bodies"][int(o.model["dofs"][d]["body"])]["qidx"])] for d in md]
        assert o.arm_limit_logit(q4) > 1.0


def test_oracle_teleports_back_to_last_valid_pose(human_data):
    """Arm swung past the valid region with a remembered valid pose: after the step the four joints sit exactly on
    the remembered pose with zero velocity (env.py:368-371)."""
    from assistive_vr_gym_b200.compiler.reset import sample_states
    from oracle.oracle import Oracle, env_to_f64
    blobs, resets = human_data
    env, var = sample_states(resets, 4, np.random.RandomState(2))
    o = Oracle(blobs[int(var[0])])
    md = [int(d) for d in o.model["header"]["mlp_dof"]]
    qi = [int(o.model["bodies"][int(o.model["dofs"][d]["body"])]["qidx"]) for d in md]
    rec = env_to_f64(env[0]).copy()
    valid = [rec[i] for i in qi]
    rec[157:161] = valid; rec[161] = 1
    rec[qi[0]] = 2.6                              # tz far inside the rejected region (logit ~ -18)
    assert o.arm_limit_logit([rec[i] for i in qi]) < -1
    o.step(rec, np.zeros(17, dtype=np.float32))
    assert np.allclose([rec[i] for i in qi], valid, atol=0.05)           # back at the valid pose (+ 4 sub-steps of motion)
    assert o.arm_limit_logit([rec[i] for i in qi]) > 0


@pytest.mark.gpu
def test_gpu_classifier_matches_oracle(human_data):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from assistive_vr_gym_b200 import capi
    from oracle.oracle import Oracle
    blobs, _ = human_data
    sim = capi.Sim(8, 0)
    for v, b in enumerate(blobs):
        sim.upload_model(v, b)
    o = Oracle(blobs[0])
    rng = np.random.RandomState(1)
    qs = rng.uniform(-3.5, 3.5, (4096, 4)).astype(np.float32)
    q_dev = torch.as_tensor(qs, device="cuda"); out = torch.zeros(4096, device="cuda")
    sim.arm_limit_logits(0, q_dev.data_ptr(), out.data_ptr(), 4096, 0)
    torch.cuda.synchronize()
    lg = out.cpu().numpy()
    ref = np.array([o.arm_limit_logit(q.astype(np.float64)) for q in qs])
    assert np.abs(lg - ref).max() < 2e-3, np.abs(lg - ref).max()
    decided = np.abs(ref) > 5e-3
    assert decided.mean() > 0.99
    assert np.array_equal(lg[decided] > 0, ref[decided] > 0)             # class (the reward-path decision) exact
    sim.close()


@pytest.mark.gpu
def test_gpu_substep_with_arm_limit_rejections(human_data):
    """One sub-step (frame_skip patched to 1) from states straddling the classifier boundary: the CUDA path must
    accept / teleport exactly where the oracle does."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from assistive_vr_gym_b200 import capi
    from assistive_vr_gym_b200.compiler.reset import sample_states
    from oracle.oracle import Oracle, env_to_f64
    from helpers import patch_blob
    blobs, resets = human_data
    pb = [patch_blob(b, header={"substeps": 1}) for b in blobs]
    nomlp = [patch_blob(b, header={"substeps": 1, "n_mlp": 0}) for b in blobs]
    n = 128
    env0, variant = sample_states(resets, n, np.random.RandomState(11))
    o0 = Oracle(pb[0])
    md = [int(d) for d in o0.model["header"]["mlp_dof"]]
    qi = [int(o0.model["bodies"][int(o0.model["dofs"][d]["body"])]["qidx"]) for d in md]
    rng = np.random.RandomState(12)
    iv = env0.view(np.int32)
    for e in range(n):
        env0[e, 157:161] = env0[e, qi]; iv[e, 161] = 1
        env0[e, qi[0]] = rng.uniform(1.6, 2.4)                           # boundary near tz ~ 2.0
        env0[e, 32 + md[0]] = rng.uniform(-1.0, 1.0)
    sim = capi.Sim(n, 0)
    for v, b in enumerate(pb):
        sim.upload_model(v, b)
    sim.set_state(env0, variant)
    obs = torch.zeros((n, 64), device="cuda"); rew = torch.zeros(n, device="cuda"); info = torch.zeros((n, 2), device="cuda")
    done = torch.zeros(n, dtype=torch.uint8, device="cuda")
    a = rng.uniform(-1, 1, (n, 17)).astype(np.float32)
    sim.step(torch.as_tensor(a, device="cuda").data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(), info.data_ptr(), 0)
    torch.cuda.synchronize()
    st = sim.get_state()
    oracles = [Oracle(b) for b in pb]; plain = [Oracle(b) for b in nomlp]
    n_rej = n_acc = 0
    for e in range(n):
        v = int(variant[e])
        pre = env_to_f64(env0[e]).copy(); plain[v].step(pre, a[e])     # pose the classifier sees (no teleport)
        logit = oracles[v].arm_limit_logit([pre[i] for i in qi])
        if abs(logit) < 0.05:
            continue                                                     # float32 vs float64 may legitimately disagree
        rec = env_to_f64(env0[e]).copy(); oracles[v].step(rec, a[e])
        rejected = logit <= 0
        n_rej += rejected; n_acc += not rejected
        assert np.abs(rec[:32] - st[e, :32]).max() < 1e-4, (e, logit)
        assert np.abs(rec[32:64] - st[e, 32:64]).max() < 2e-3, (e, logit)
        assert np.abs(rec[157:161] - st[e, 157:161]).max() < 1e-4        # remembered pose
        if rejected:
            assert np.array_equal(st[e, qi], env0[e, 157:161]) and not st[e, [32 + d for d in md]].any()
    assert n_rej >= 10 and n_acc >= 10, (n_rej, n_acc)
    sim.close()
