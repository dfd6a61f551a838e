"""Size-independent properties of the CUDA path at BASELINE sizes, ragged sizes, and the host-buffer entry point."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _run(torch, n, steps, seed=1, replicate=False):
    from assistive_vr_gym_b200 import make
    env = make("ScratchItchJaco-v0", num_envs=n, device=0, seed=seed)
    env.reset()
    if replicate:                                    # every env = env 0
        st = env.get_state()
        env.set_state(np.repeat(st[:1], n, axis=0), np.repeat(env.variants[:1], n))
    gen = torch.Generator(device="cuda"); gen.manual_seed(0)
    outs = []
    for t in range(steps):
        a = torch.rand((1 if replicate else n, 7), device="cuda", generator=gen) * 2 - 1
        if replicate:
            a = a.expand(n, 7).contiguous()
        obs, rew, done, info = env.step(a)
        outs.append((obs.clone(), rew.clone()))
    st = env.get_state()
    env.close()
    return outs, st


@pytest.mark.parametrize("n", [1, 3, 33, 130])
def test_ragged_batch_sizes(torch_cuda, n):
    """Batch sizes that are not multiples of the warps-per-block: every env is stepped, none twice."""
    outs, st = _run(torch_cuda, n, 2)
    assert st.shape[0] == n and np.all(st.view(np.int32)[:, 152] == 2)           # iteration counter
    assert np.isfinite(st).all()


def test_determinism_and_replication_at_scale(torch_cuda):
    """32768 envs (config 5's per-box size): identical envs give bit-identical outputs; reruns are bit-identical."""
    torch = torch_cuda
    n = 32768
    o1, s1 = _run(torch, n, 3, replicate=True)
    for obs, rew in o1:
        assert torch.equal(obs, obs[:1].expand_as(obs)) and torch.equal(rew, rew[:1].expand_as(rew))
    o2, s2 = _run(torch, n, 3, replicate=True)
    assert np.array_equal(s1, s2)


def test_state_invariants_at_scale(torch_cuda):
    torch = torch_cuda
    n = 65536
    outs, st = _run(torch, n, 5, seed=2)
    iv = st.view(np.int32)
    assert np.isfinite(st[:, :64]).all() and np.isfinite(outs[-1][0].cpu().numpy()).all()
    assert np.abs(np.linalg.norm(st[:, 20:24], axis=1) - 1).max() < 1e-4          # tool quaternion stays unit
    assert np.all(iv[:, 166] & 3 == 0)                                            # no contact / row overflow
    assert np.all(iv[:, 152] == 5) and np.all(iv[:, 167] <= 5 * 50)
    # human joints never leave their (scaled) limits: enforce_hard_human_joint_limits, env.py:389-410
    from assistive_vr_gym_b200.envs import load_env_data
    _, resets = load_env_data("ScratchItchJaco.npz")
    rd = resets[0]
    hq = st[:, rd["hum_qidx"]]; ls = st[:, 97][:, None]
    assert np.all(hq >= rd["hum_lower"][None] * ls - 1e-5) and np.all(hq <= rd["hum_upper"][None] * ls + 1e-5)
    # the weld keeps the tool at the end effector
    obs = outs[-1][0].cpu().numpy()
    assert np.percentile(np.linalg.norm(obs[:, 0:3], axis=1), 99) < 1.2


@pytest.mark.parametrize("env_id,n", [("ScratchItchJaco-v0", 257), ("BedBathingPR2-v0", 130), ("ScratchItchPR2Human-v0", 66), ("ScratchItchJaco-v0", 40000), ("ScratchItchJaco-v0", 140000)])
def test_step_host_equals_device_step(torch_cuda, env_id, n):
    """The host-buffer entry point (pinned staging, two chunks on two streams for 16384 .. 262143 envs) returns bit for bit what the
    device-resident step returns (itself two half batches on two streams for 2048 .. 131071 envs, one sequence above: at 140000 envs
    the two partition the batch differently, and both use the handle's large-batch collide instance)."""
    torch = torch_cuda
    from assistive_vr_gym_b200 import make
    e1 = make(env_id, num_envs=n, device=0, seed=4); e1.reset()
    e2 = make(env_id, num_envs=n, device=0, seed=4); e2.reset()
    a = np.random.RandomState(0).uniform(-1, 1, (3, n, e1.sim.n_actions)).astype(np.float32)
    for t in range(3):
        o1, r1, d1, i1 = e1.step(torch.as_tensor(a[t], device="cuda"))
        o2, r2, d2, i2 = e2.step_host(a[t])
        assert np.array_equal(o1.cpu().numpy(), o2) and np.array_equal(r1.cpu().numpy(), r2)
        assert np.array_equal(i1["task_success"].cpu().numpy(), i2["task_success"])
    e1.close(); e2.close()


@pytest.mark.parametrize("env_id,n", [("ScratchItchJaco-v0", 5003), ("BedBathingPR2-v0", 1026), ("FeedingJaco-v0", 70)])
def test_step_host_one_chunk_with_overlapped_result_copies(torch_cuda, monkeypatch, env_id, n):
    """Large batches step as ONE sequence and run the epilogue range by range, each range's device->host copies on the second
    stream (avg_step_host, from 262144 envs; forced here at a small batch through AVG_CHUNKS / AVG_TAIL_RANGES): bit for bit the
    device-resident step, over pinned and over pageable caller buffers."""
    torch = torch_cuda
    from assistive_vr_gym_b200 import make
    monkeypatch.setenv("AVG_CHUNKS", "1"); monkeypatch.setenv("AVG_TAIL_RANGES", "4")
    e1 = make(env_id, num_envs=n, device=0, seed=4); e1.reset()
    e2 = make(env_id, num_envs=n, device=0, seed=4); e2.reset()
    a = np.random.RandomState(0).uniform(-1, 1, (4, n, e1.sim.n_actions)).astype(np.float32)
    pin = e2.pinned_actions()
    for t in range(4):
        o1, r1, d1, i1 = e1.step(torch.as_tensor(a[t], device="cuda"))
        if t % 2 == 0:
            o2, r2, d2, i2 = e2.step_host(a[t])                  # pageable actions: staged
        else:
            pin[...] = a[t]; o2, r2, d2, i2 = e2.step_host(pin)  # page-locked actions: in place
        assert np.array_equal(o1.cpu().numpy(), o2) and np.array_equal(r1.cpu().numpy(), r2)
        assert np.array_equal(i1["total_force_on_human"].cpu().numpy(), i2["total_force_on_human"])
        assert np.array_equal(i1["task_success"].cpu().numpy(), i2["task_success"])
    e1.close(); e2.close()


def test_reference_shaped_single_env(torch_cuda):
    """examples/random_actions.py shape: make -> reset -> step(action_space.sample()) with the reference's types."""
    from assistive_vr_gym_b200 import make
    env = make("ScratchItchJaco-v0")
    obs = env.reset()
    assert obs.dtype == np.float64 and obs.shape == (30,)
    obs, reward, done, info = env.step(env.action_space.sample())
    assert obs.shape == (30,) and isinstance(reward, float) and done is False
    assert set(info) >= {"total_force_on_human", "task_success", "action_robot_len", "action_human_len", "obs_robot_len", "obs_human_len"}
    assert info["action_robot_len"] == 7 and info["obs_robot_len"] == 30 and info["task_success"] in (0, 1)
    env.close()


def test_time_limit(torch_cuda):
    torch = torch_cuda
    from assistive_vr_gym_b200 import make
    env = make("ScratchItchJaco-v0", num_envs=4)
    env.reset()
    a = torch.zeros((4, 7), device="cuda")
    for t in range(200):
        obs, rew, done, info = env.step(a)
        assert bool(done.any()) == (t == 199)        # TimeLimit(200), __init__.py:21; the env itself never terminates
    env.close()


def test_unknown_ids():
    from assistive_vr_gym_b200 import make
    with pytest.raises(NotImplementedError):
        make("ScratchItchVRJaco-v0", num_envs=1)                          # registered by the reference, out of scope (VR / headset ids)
    with pytest.raises(KeyError):
        make("NoSuchEnv-v0", num_envs=1)


def test_separating_axis_cache_only_skips_work(torch_cuda):
    """The collide kernel's temporal cache (pair -> last separating direction) may reject a candidate early but must
    never change a contact: trajectories with the cache disabled (AVG_DBG bit 4) are bit-identical."""
    torch = torch_cuda
    from assistive_vr_gym_b200 import make
    n, T = 4096, 12
    rng = np.random.RandomState(5)
    acts = rng.uniform(-1, 1, (T, n, 7)).astype(np.float32)
    out = []
    for dbg in ("16", "0"):
        os.environ["AVG_DBG"] = dbg
        try:
            env = make("ScratchItchJaco-v0", num_envs=n, device=0, seed=21)
            env.sim.enable_debug(True)
            env.reset()
            ncont = 0
            for t in range(T):
                env.step(torch.as_tensor(acts[t], device="cuda"))
                ncont += int(env.sim.get_contacts()[1].sum())
            out.append((env.get_state().copy(), env.reward.cpu().numpy().copy(), ncont))
            env.close()
        finally:
            os.environ.pop("AVG_DBG", None)
    assert out[0][2] == out[1][2] and out[0][2] > 100            # contacts happened and are the same in number
    assert np.array_equal(out[0][0][:, :64], out[1][0][:, :64])  # q, qd bit-identical
    assert np.array_equal(out[0][1], out[1][1])


@pytest.mark.parametrize("chunks", [1, 2, 4])
def test_step_is_cuda_graph_capturable(torch_cuda, chunks, monkeypatch):
    """the step is a fixed launch sequence with no hidden synchronisation, so it can be captured once and
    replayed; replays are bit-identical to eager steps.  chunks = 2: the two-stream form large batches use (half batches on
    the caller's stream and on the handle's second stream, joined by events) is captured as a forked graph and equals the
    single-stream sequence bit for bit."""
    torch = torch_cuda
    from assistive_vr_gym_b200 import make
    n = 4096
    acts = torch.rand((6, n, 7), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)) * 2 - 1
    monkeypatch.setenv("AVG_STEP_CHUNKS", "1")
    ref = make("ScratchItchJaco-v0", num_envs=n, device=0, seed=31); ref.reset()
    monkeypatch.setenv("AVG_STEP_CHUNKS", str(chunks))               # read when the handle is created
    env = make("ScratchItchJaco-v0", num_envs=n, device=0, seed=31); env.reset()
    for t in range(2):                                         # warm-up outside the capture (function attributes, lazy state)
        ref.step(acts[t]); env.step(acts[t])
    static_act = torch.zeros((n, 7), device="cuda")
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        static_act.copy_(acts[2])
        side.synchronize()
        g.capture_begin()
        env.sim.step(static_act.data_ptr(), env.obs.data_ptr(), env.reward.data_ptr(), env.done_dev.data_ptr(), env.info_dev.data_ptr(),
                     int(torch.cuda.current_stream().cuda_stream))
        g.capture_end()
    for t in range(2, 6):
        static_act.copy_(acts[t])
        g.replay()
        ref.step(acts[t])
    torch.cuda.synchronize()
    assert torch.equal(env.obs, ref.obs) and torch.equal(env.reward, ref.reward)
    assert np.array_equal(env.get_state()[:, :64], ref.get_state()[:, :64])
    env.close(); ref.close()


def test_auto_reset_on_the_device(torch_cuda):
    """auto_reset=True after a device reset: step 200 reports done (TimeLimit), hands back the terminal observation in info
    and the first observation of the next episode as `obs`, restarted on the GPU (next episode index: different draws),
    and the batch keeps stepping."""
    torch = torch_cuda
    from assistive_vr_gym_b200 import make
    n = 256
    env = make("ScratchItchJaco-v0", num_envs=n, device=0, seed=4, auto_reset=True, device_ik=True)
    first = env.reset_device(seed=11).clone()
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    for t in range(200):
        obs, rew, done, info = env.step(torch.rand((n, 7), device="cuda", generator=g) * 2 - 1)
        assert bool(done.all()) == (t == 199)
    assert "terminal_observation" in info and info["terminal_observation"].shape == obs.shape
    assert env.elapsed == 0
    st = env.get_state()
    assert np.all(st.view(np.int32)[:, 152] == 0)                 # AVG_E_ITERATION: a fresh episode in every environment
    assert not torch.equal(obs, first) and not torch.equal(obs, info["terminal_observation"])
    obs2, rew2, done2, info2 = env.step(torch.zeros((n, 7), device="cuda"))
    assert not bool(done2.any()) and torch.isfinite(obs2).all()
    env.close()


def test_mixed_task_batch(torch_cuda):
    """One sub-batch per id on separate streams (mixed.py): every sub-batch returns exactly what the same id returns when
    stepped alone (handles are independent), with the right shapes per id."""
    torch = torch_cuda
    from assistive_vr_gym_b200 import make
    from assistive_vr_gym_b200.mixed import MixedBatch
    ids = ["ScratchItchJaco-v0", "BedBathingPR2-v0", "ScratchItchPR2Human-v0", "BedBathingJacoHuman-v0"]
    n = 96
    mb = MixedBatch(ids, envs_per_id=n, device=0, seed=50)
    mb.reset_device(seed=8)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    acts = [mb.sample_actions(g) for _ in range(4)]
    for a in acts:
        out = mb.step(a)
    torch.cuda.synchronize()
    for k, i in enumerate(ids):
        solo = make(i, num_envs=n, device=0, seed=50 + k)
        solo.reset_device(seed=8)
        for a in acts:
            o, r, d, info = solo.step(a[i])
        assert out[i][0].shape == (n, solo.sim.n_obs) and acts[0][i].shape == (n, solo.sim.n_actions)
        assert torch.equal(out[i][0], o) and torch.equal(out[i][1], r)
        solo.close()
    mb.close()


def test_staggered_episodes_with_device_time_limit(torch_cuda):
    """auto_reset="device": gym's TimeLimit(200) is evaluated per environment on the device (avg_set_time_limit), and the
    environments whose done byte is set restart inside step() with that byte array as avg_reset's mask.  Half of the batch
    is restarted by hand after 50 steps, so the halves end their episodes 50 steps apart."""
    torch = torch_cuda
    from assistive_vr_gym_b200 import make
    n = 64
    env = make("ScratchItchJaco-v0", num_envs=n, device=0, seed=6, auto_reset="device", device_ik=True)
    env.reset_device(seed=21)
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    first_half = torch.zeros(n, dtype=torch.uint8, device="cuda"); first_half[: n // 2] = 1
    ends = []
    for t in range(1, 261):
        obs, rew, done, info = env.step(torch.rand((n, 7), device="cuda", generator=g) * 2 - 1)
        d = done.cpu().numpy()
        if d.any():
            ends.append((t, d.copy()))
            to = info["terminal_observation"]
            assert not torch.equal(to[done], obs[done])                       # restarted rows carry the next episode's first observation
            assert torch.equal(to[~done], obs[~done])
            it = env.get_state().view(np.int32)[:, 152]
            assert np.all(it[d] == 0) and np.all(it[~d] > 0)                  # AVG_E_ITERATION: fresh episodes exactly where done was set
        if t == 50:
            env.reset_device(mask=first_half)
    assert [e[0] for e in ends] == [200, 250]
    assert ends[0][1][n // 2:].all() and not ends[0][1][: n // 2].any()       # second half: 200 steps since the common reset
    assert ends[1][1][: n // 2].all() and not ends[1][1][n // 2:].any()       # first half: 200 steps since its own restart
    env.close()


@pytest.mark.parametrize("env_id", ["ScratchItchJaco-v0", "BedBathingPR2-v0"])
def test_cuda_graph_mode_equals_eager_steps(torch_cuda, env_id):
    """make(..., cuda_graph=True): step() replays one captured graph of the launch sequence; observations, rewards, done bytes
    and the state records equal those of eager steps bit for bit, for caller-owned action tensors that change every step."""
    torch = torch_cuda
    from assistive_vr_gym_b200 import make
    n = 1024
    ref = make(env_id, num_envs=n, device=0, seed=41); ref.reset()
    env = make(env_id, num_envs=n, device=0, seed=41, cuda_graph=True); env.set_state(ref.get_state(), ref.variants)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    for t in range(7):
        a = torch.rand((n, env.sim.n_actions), device="cuda", generator=g) * 2 - 1
        o0, r0, d0, i0 = ref.step(a)
        o1, r1, d1, i1 = env.step(a.clone())
        assert torch.equal(o0, o1) and torch.equal(r0, r1) and torch.equal(d0, d1)
        assert torch.equal(i0["total_force_on_human"], i1["total_force_on_human"])
    assert "step" in env._graphs and env._eager_steps == 1            # one eager step, then replays
    # bit patterns, not float values: the record holds integer words too (the alive-target bitmap of BedBathing reads as NaN)
    assert np.array_equal(env.get_state().view(np.uint32), ref.get_state().view(np.uint32))
    env.close(); ref.close()


def test_cuda_graph_rollout_with_device_auto_reset(torch_cuda):
    """The graph of rollout() holds policy -> step -> masked device restart (auto_reset="device"); with a 5-step time limit the
    restarts happen inside the replays and the run equals the eager one bit for bit."""
    torch = torch_cuda
    from assistive_vr_gym_b200 import make
    from assistive_vr_gym_b200.policy import pack_policy, synthetic_policy
    n = 512
    _, arrs = synthetic_policy(30, 7, seed=3)
    envs = []
    for graphed in (False, True):
        e = make("ScratchItchJaco-v0", num_envs=n, device=0, seed=8, auto_reset="device", device_ik=True, cuda_graph=graphed)
        e.sim.set_time_limit(5)
        e.set_policy(pack_policy(**arrs))
        e.reset_device(seed=33)
        envs.append(e)
    ends = []
    for t in range(1, 13):
        o0, r0, d0, i0 = envs[0].rollout(1)
        o1, r1, d1, i1 = envs[1].rollout(1)
        assert torch.equal(o0, o1) and torch.equal(r0, r1) and torch.equal(d0, d1)
        assert torch.equal(i0["terminal_observation"], i1["terminal_observation"])
        if bool(d1.all()):
            ends.append(t)
    assert ends == [5, 10] and "rollout" in envs[1]._graphs
    assert np.array_equal(envs[0].get_state().view(np.uint32), envs[1].get_state().view(np.uint32))
    for e in envs:
        e.close()


@pytest.mark.parametrize("env_id,n", [("ScratchItchJaco-v0", 1500), ("BedBathingPR2-v0", 700), ("ScratchItchJacoHuman-v0", 300)])
def test_fused_dynamics_solve_kernel_is_bit_identical(torch_cuda, env_id, n):
    """Small launches run dynamics + solve of a sub-step as one kernel (avg_dynsolve_kernel, AVG_FUSE): the same two bodies, so the
    environment records after a seeded roll are bit for bit those of the two-kernel pipeline.  The switch is read once per process,
    hence the two sub-processes."""
    import re, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sha = []
    for fuse in ("0", "1"):
        out = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_state_hash.py"), env_id, str(n), "12"], capture_output=True, text=True,
                             env=dict(os.environ, AVG_FUSE=fuse), timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        sha.append(re.search(r"state sha1 ([0-9a-f]{40})", out.stdout).group(1))
    assert sha[0] == sha[1]


@pytest.mark.parametrize("n,expect", [(1000, 17), (3000, 34), (8192, 44), (140000, 22)])
def test_launches_per_env_step(torch_cuda, n, expect):
    """What one env-step launches (ScratchItchJaco-v0): 22 kernels as one sequence from 131072 environments; two stream halves of 22
    below that; launches of <= 2048 environments fuse dynamics + solve (17 per sequence)."""
    torch = torch_cuda
    from assistive_vr_gym_b200 import make
    env = make("ScratchItchJaco-v0", num_envs=n, device=0, seed=2); env.reset()
    a = torch.zeros((n, 7), device="cuda")
    env.step(a)
    c0 = env.sim.launch_count
    env.step(a)
    assert env.sim.launch_count - c0 == expect
    env.close()
