"""Feeding / Drinking (reference feeding.py, drinking.py): compiled model, CPU oracle, and the CUDA path against the oracle.

CPU tests pin the oracle against plain-numpy restatements of the in-tree reference formulas (observation layout,
get_food_rewards / get_water_rewards, points_in_cylinder, the tilt term) and against physical sanity (the food settles into
the spoon).  GPU tests (`-m gpu`) compare the CUDA kernels with the oracle from identical states: particle event indices
(eaten / spilled / hit) bit-exact, joint positions, particle positions and rewards within the tolerances written below.
Parity with PyBullet itself is unpinned (no PyBullet in this environment).
"""
import numpy as np
import pytest

from assistive_vr_gym_b200.compiler.blob import read_blob
from assistive_vr_gym_b200.compiler.reset_fd import sample_states_fd
from assistive_vr_gym_b200.envs import REGISTRY, load_env_data
from helpers import quat_rot

P_POS, P_VEL, P_ALIVE, P_HIT, P_TOUCH_H, P_TOUCH_S, P_EV_EAT, P_EV_SPILL, P_EV_HIT = 0, 192, 576, 578, 580, 582, 584, 586, 588
E_TASK_SUCCESS, E_TARGET_POS = 153, 162
F_TOOL, F_TORSO, F_CHEST, F_HEAD = 0, 3, 4, 8
FD_IDS = [k for k in REGISTRY if k.startswith(("Feeding", "Drinking"))]


def _oracle_env(env_id, seed=0, gender=0):
    from oracle.oracle import Oracle, env_to_f64, part_to_f64
    blobs, resets = load_env_data(REGISTRY[env_id]["data"])
    env, part, var = sample_states_fd(resets, 1, np.random.RandomState(seed), genders=np.array([gender]))
    o = Oracle(blobs[int(var[0])])
    return o, env_to_f64(env[0]).copy(), part_to_f64(part[0]).copy()


def test_registry_has_every_feeding_drinking_id():
    """reference __init__.py registers <Task>{PR2,Jaco}[Human]; BASELINE.json adds Sawyer and Baxter (build-defined ids)."""
    for t in ("Feeding", "Drinking"):
        for r in ("Jaco", "PR2", "Sawyer", "Baxter"):
            for v in ("", "Human"):
                assert f"{t}{r}{v}-v0" in REGISTRY


@pytest.mark.parametrize("env_id,n_part,n_obs,n_act", [("FeedingJaco-v0", 8, 25, 7), ("DrinkingJaco-v0", 64, 25, 7),
                                                       ("FeedingJacoHuman-v0", 8, 48, 11), ("DrinkingJacoHuman-v0", 64, 48, 11)])
def test_blob_layout(env_id, n_part, n_obs, n_act):
    """feeding.py:18 (7 + 4 actions, 25 + 23 observations), :289 (numSubSteps 2, 10 iterations), :300 (8 spheres) / drinking.py:301
    (64); the tool is one compound of VHACD hulls (spoon 64, cup 68), the head another, Feeding adds the bowl (70, env-static)."""
    blobs, resets = load_env_data(REGISTRY[env_id]["data"])
    m = read_blob(blobs[0]); h = m["header"]
    assert int(h["n_particle"]) == n_part and int(h["n_internal"]) == 2 and int(h["solver_iters"]) == 10 and abs(float(h["dt"]) - 0.01) < 1e-9
    assert int(h["n_obs_robot"] + h["n_obs_human"]) == n_obs and int(h["n_action_robot"] + h["n_action_human"]) == n_act
    sh = m["shapes"]
    comp = [i for i in range(int(h["n_shape"])) if sh[i]["type"] == 6]
    feeding = env_id.startswith("Feeding")
    assert len(comp) == (3 if feeding else 2) and all(i < int(h["n_mshape"]) for i in comp)
    counts = sorted(int(sh[i]["vert_cnt"]) for i in comp)
    assert counts == ([8, 64, 70] if feeding else [8, 68])                       # male head: 8 hulls
    assert int(h["n_ebody"]) == (1 if feeding else 0)
    assert int(h["pshape"]) == int(h["n_shape"] + h["n_cshape"]) and abs(float(sh[int(h["pshape"])]["radius"]) - 0.005) < 1e-9
    for i in comp:                                                              # children carry their compound and a box inside its box
        first, cnt = int(sh[i]["vert_off"]), int(sh[i]["vert_cnt"])
        assert all(int(sh[c]["pad"][0]) == i and sh[c]["type"] == 4 and sh[c]["body"] == sh[i]["body"] for c in range(first, first + cnt))
        ca = m["caabb"][first - int(h["n_shape"]):first - int(h["n_shape"]) + cnt]
        lo = (ca[:, 0:3] - ca[:, 4:7]).min(0); hi = (ca[:, 0:3] + ca[:, 4:7]).max(0)
        assert np.all(lo >= sh[i]["aabb_c"] - sh[i]["aabb_h"] - 1e-5) and np.all(hi <= sh[i]["aabb_c"] + sh[i]["aabb_h"] + 1e-5)
    assert float(h["task_f"][4]) == pytest.approx(0.75 * n_part)                # task_success_threshold * count, feeding.py:76


def test_struct_sizes_match_the_oracle():
    from oracle.oracle import Oracle
    from assistive_vr_gym_b200.compiler.blob import HEADER_DT
    from assistive_vr_gym_b200.compiler.reset import RESET_TABLE_DT
    blobs, _ = load_env_data("FeedingJaco.npz")
    sizes = Oracle(blobs[0]).sizes()
    assert sizes[0] == HEADER_DT.itemsize and sizes[6] == RESET_TABLE_DT.itemsize


@pytest.mark.parametrize("env_id", ["FeedingJaco-v0", "DrinkingJaco-v0"])
def test_oracle_settle_drops_particles_into_the_tool(env_id):
    """feeding.py:318-320: after the 100 settle steps the particles rest in the spoon / cup: all alive, nearly at rest, below
    their creation height and still within 5 cm (spoon) / 8 cm (cup) of the tool."""
    o, env, part = _oracle_env(env_id, seed=3)
    npart = o.n_particle
    z0 = part[P_POS + 128:P_POS + 128 + npart].copy()
    o.settle(env, part, 100)
    h = o.model["header"]
    tq = int(o.model["bodies"][int(h["tool_body"])]["qidx"])
    tool = env[tq:tq + 3]
    x = np.stack([part[P_POS + 64 * c:P_POS + 64 * c + npart] for c in range(3)], axis=1)
    v = np.stack([part[P_VEL + 64 * c:P_VEL + 64 * c + npart] for c in range(3)], axis=1)
    assert np.all(x[:, 2] < z0 + 1e-6) and np.all(x[:, 2] > tool[2] - 0.02)
    assert np.linalg.norm(x - tool, axis=1).max() < (0.05 if env_id.startswith("Feeding") else 0.13)
    assert np.abs(v).max() < 0.05
    assert int(part[590]) > npart and int(part[591]) == 0                        # resting contacts, no overflow


def _obs_numpy(o, env, tool_force, robot_force, human_control):
    """feeding.py:123-142 restated with numpy from the oracle's frames."""
    torso = o.frame(env, F_TORSO)[:3]; tool = o.frame(env, F_TOOL); head = o.frame(env, F_HEAD); chest = o.frame(env, F_CHEST)[:3]
    mouth_pos = np.asarray(o.model["header"]["task_f"][19:22], dtype=np.float64)
    mouth = head[:3] + quat_rot(head[3:], mouth_pos)
    dofs = o.model["dofs"]; bodies = o.model["bodies"]
    arm = [env[int(bodies[int(d["body"])]["qidx"])] for d in dofs[:int(o.model["header"]["n_jdof"])] if 0 <= d["action"] < 7]
    robot_obs = np.concatenate([tool[:3] - torso, tool[3:], tool[:3] - mouth, arm, head[:3] - torso, head[3:], [tool_force]])
    if not human_control:
        return robot_obs, mouth
    hq = np.zeros(4)
    for d in dofs[:int(o.model["header"]["n_jdof"])]:
        if d["human_slot"] >= 0:
            hq[int(d["human_slot"])] = env[int(bodies[int(d["body"])]["qidx"])]
    human_obs = np.concatenate([tool[:3] - chest, tool[3:], tool[:3] - mouth, hq, head[:3] - chest, head[3:], [robot_force, tool_force]])
    return np.concatenate([robot_obs, human_obs]), mouth


@pytest.mark.parametrize("env_id", ["FeedingJaco-v0", "FeedingJacoHuman-v0", "DrinkingJacoHuman-v0"])
def test_oracle_observation_layout(env_id):
    o, env, part = _oracle_env(env_id, seed=5)
    obs = o.reset_obs(env)
    ref, mouth = _obs_numpy(o, env, 0.0, 0.0, env_id.endswith("Human-v0"))
    assert obs.shape == ref.shape and np.abs(obs - ref).max() < 1e-12
    rng = np.random.RandomState(1)
    a = rng.uniform(-1, 1, o.n_act).astype(np.float32)
    obs, rew, info, cont = o.step(env, a, part)
    ref, mouth = _obs_numpy(o, env, info[3], info[2], env_id.endswith("Human-v0"))
    assert np.abs(obs - ref).max() < 1e-12
    assert np.abs(env[E_TARGET_POS:E_TARGET_POS + 3] - mouth).max() < 1e-12      # update_targets, feeding.py:345-349


def _set_particle(part, p, x, v=(0, 0, 0)):
    for c in range(3):
        part[P_POS + 64 * c + p] = x[c]; part[P_VEL + 64 * c + p] = v[c]


def test_oracle_food_rewards():
    """feeding.py:92-121 against a numpy restatement: a particle at the mouth is eaten (+20, counted, its speed penalised), one
    below z = 0.5 is spilled (-5), the others stay; reward = distance + action + food + preferences (feeding.py:71)."""
    from oracle.oracle import part_masks
    o, env, part = _oracle_env("FeedingJaco-v0", seed=7)
    o.settle(env, part, 20)
    a = np.zeros(7, dtype=np.float32)
    obs, rew0, info0, _ = o.step(env, a, part)                                  # a quiet step: no events
    assert part_masks(part, P_ALIVE) == 0xff and info0[6] == 0.0
    mouth = env[E_TARGET_POS:E_TARGET_POS + 3].copy()
    head = o.frame(env, F_HEAD)
    _set_particle(part, 2, mouth + 0.014 * quat_rot(head[3:], [0, -1, 0]), v=(0.0, 0.0, 0.5396))   # just in front of the lips, thrown up so that it is back there after 10 x 0.01 s of gravity
    _set_particle(part, 5, np.array([0.3, -0.3, 0.45]))                         # below z = 0.5
    success0 = env[E_TASK_SUCCESS]
    obs, rew, info, _ = o.step(env, a, part)
    assert part_masks(part, P_EV_EAT) == 1 << 2 and part_masks(part, P_EV_SPILL) == 1 << 5 and part_masks(part, P_EV_HIT) == 0
    assert part_masks(part, P_ALIVE) == 0xff & ~((1 << 2) | (1 << 5))
    assert env[E_TASK_SUCCESS] == success0 + 1
    assert info[6] == pytest.approx(20.0 - 5.0)
    speed = np.linalg.norm([part[P_VEL + 64 * c + 2] for c in range(3)])
    tool = o.frame(env, F_TOOL)[:3]
    tq = int(o.model["bodies"][int(o.model["header"]["tool_body"])]["dof"])
    ee_vel = np.linalg.norm(env[32 + tq:32 + tq + 3])
    pref = 0.25 * (-ee_vel) + 0.01 * (-info[2]) + 0.05 * (0.0 if info[3] < 10 else -info[3]) + 1.0 * 0.0 + 1.0 * (-speed)   # env.py:412-448
    expected = 1.0 * (-np.linalg.norm(mouth_after(o, env) - tool)) + 0.01 * 0.0 + 1.0 * 15.0 + pref
    assert rew == pytest.approx(expected, abs=1e-9)
    obs, rew, info, _ = o.step(env, a, part)                                    # removed particles are not counted again
    assert part_masks(part, P_EV_EAT) == 0 and part_masks(part, P_EV_SPILL) == 0 and info[6] == 0.0


def mouth_after(o, env):
    head = o.frame(env, F_HEAD)
    return head[:3] + quat_rot(head[3:], np.asarray(o.model["header"]["task_f"][19:22], dtype=np.float64))


def test_oracle_water_rewards_and_cylinder():
    """drinking.py:95-136 + util.py:107-110: water inside the cup's cylinder is not looked at; outside it, near the mouth counts
    +10 (speed read after the teleport = 0), below z = 0.5 counts -1; the tilt term follows drinking.py:71-72."""
    from oracle.oracle import part_masks
    o, env, part = _oracle_env("DrinkingJaco-v0", seed=9)
    o.settle(env, part, 30)
    a = np.zeros(7, dtype=np.float32)
    obs, rew0, info0, _ = o.step(env, a, part)
    full = (1 << 64) - 1
    assert part_masks(part, P_ALIVE) == full                                    # everything still inside the cup
    tool = o.frame(env, F_TOOL)
    # the reference's cylinder test on the settled water: all inside
    rx = np.array([np.sin(np.pi / 4), 0, 0, np.cos(np.pi / 4)])

    def qmul(a_, b_):
        x1, y1, z1, w1 = a_; x2, y2, z2, w2 = b_
        return np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                         w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])
    cup_p = tool[:3] + quat_rot(tool[3:], [0, 0.06, 0]); cup_q = qmul(tool[3:], rx)
    top = cup_p + quat_rot(cup_q, [0, 0, -0.055]); bottom = cup_p + quat_rot(cup_q, [0, 0, 0.07])
    x = np.stack([part[P_POS + 64 * c:P_POS + 64 * c + 64] for c in range(3)], axis=1)
    vec = bottom - top
    inside = ((x - top) @ vec >= 0) & ((x - bottom) @ vec <= 0) & (np.linalg.norm(np.cross(x - top, vec), axis=1) <= 0.05 * np.linalg.norm(vec))
    assert inside.all()
    mouth = env[E_TARGET_POS:E_TARGET_POS + 3].copy()
    head = o.frame(env, F_HEAD)
    _set_particle(part, 40, mouth + 0.014 * quat_rot(head[3:], [0, -1, 0]), v=(0.0, 0.0, 0.5396))  # in front of the lips, back there after 10 x 0.01 s of gravity
    _set_particle(part, 41, np.array([0.4, -0.4, 0.3]))
    s0 = env[E_TASK_SUCCESS]
    obs, rew, info, _ = o.step(env, a, part)
    assert part_masks(part, P_EV_EAT) == 1 << 40 and part_masks(part, P_EV_SPILL) == 1 << 41
    assert part_masks(part, P_ALIVE) == full & ~((1 << 40) | (1 << 41)) and env[E_TASK_SUCCESS] == s0 + 1
    # reward_water = +10 - 1; tilt = -|roll + pi/2| (Jaco) with roll of cup_q; info[6] = water + 0.1 * tilt
    tool = o.frame(env, F_TOOL); cq = qmul(tool[3:], rx)
    sarg = -2 * (cq[0] * cq[2] - cq[3] * cq[1])
    roll = 0.0 if abs(sarg) >= 0.99999 else np.arctan2(2 * (cq[1] * cq[2] + cq[3] * cq[0]), cq[3] ** 2 - cq[0] ** 2 - cq[1] ** 2 + cq[2] ** 2)
    assert info[6] == pytest.approx(9.0 + 0.1 * (-abs(roll + np.pi / 2)), abs=1e-7)           # cup_tilt_weight is stored as float32


def test_oracle_frozen_head_and_tremor():
    """feeding.py:244: without human control or tremor the head chain is static (mass 0 -> no motion); with a tremor the head's
    motor targets follow target +- tremor (env.py:330-332) and the head moves."""
    o, env, part = _oracle_env("FeedingJaco-v0", seed=11)
    dofs, bodies = o.model["dofs"], o.model["bodies"]
    head_q = [int(bodies[int(d["body"])]["qidx"]) for d in dofs[:int(o.model["header"]["n_jdof"])] if d["human_slot"] >= 0]
    assert len(head_q) == 4
    env[99] = 0.0; env[100:104] = 0.0; env[175] = float(int(o.model["header"]["head_frozen_mask"]))
    q0 = env[head_q].copy()
    a = np.random.RandomState(0).uniform(-1, 1, 7).astype(np.float32)
    for _ in range(3):
        o.step(env, a, part)
    assert np.abs(env[head_q] - q0).max() < 1e-6                                # (a head angle clipped to its scaled limit may be re-clamped in float64)
    env[99] = 1.0; env[100:104] = np.deg2rad([15, -15, 15, -15]); env[175] = 0.0
    for _ in range(3):
        o.step(env, a, part)
    assert np.abs(env[head_q] - q0).max() > 1e-3


# ------------------------------------------------------------------------------------------------------------------ GPU
def _gpu_env(env_id, n, seed):
    from assistive_vr_gym_b200 import make
    from oracle.oracle import Oracle, env_to_f64, part_to_f64
    env = make(env_id, num_envs=n, device=0, seed=seed)
    obs0 = env.reset().cpu().numpy().copy()
    state, part = env.get_state(), env.get_particles()
    oracles = [Oracle(b) for b in env.blobs]
    recs = [env_to_f64(state[e]).copy() for e in range(n)]
    parts = [part_to_f64(part[e]).copy() for e in range(n)]
    return env, obs0, oracles, recs, parts


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", ["FeedingJaco-v0", "DrinkingJaco-v0", "FeedingPR2Human-v0", "FeedingSawyer-v0", "DrinkingBaxter-v0",
                                    "FeedingJacoNew-v0", "DrinkingPR2New-v0"])
def test_gpu_reset_and_trajectory_vs_oracle(env_id):
    """Device reset (draws, IK, particle grid, 100 settle steps) leaves the food in the tool; from that state the CUDA step and the
    oracle agree over 3 env-steps (= 30 internal steps with ~30 / ~250 particle contacts each): reset observation 1e-5, joint
    positions 1e-4, particle positions 2e-3 (resting contacts amplify float32 noise), events bit-exact, reward 2e-3."""
    import torch
    from oracle.oracle import env_to_f64, part_to_f64, part_masks
    n = 6
    env, obs0, oracles, recs, parts = _gpu_env(env_id, n, seed=13)
    npart = env.sim.n_particles
    part0 = env.get_particles()
    assert np.all(part0.view(np.uint32)[:, P_ALIVE] == (0xff if npart == 8 else 0xffffffff))
    for e in range(n):
        o = oracles[int(env.variants[e])]
        assert np.abs(o.reset_obs(recs[e]) - obs0[e]).max() < 1e-5
        assert np.median(np.abs(part0[e, P_VEL:P_VEL + 192]).reshape(3, 64)[:, :npart].max(0)) < 0.1   # at rest after the settle (a start pose within the IK tolerance may let one roll off)
    rng = np.random.RandomState(2)
    env.sim.enable_debug(True)
    for s in range(3):
        a = rng.uniform(-1, 1, (n, env.sim.n_actions)).astype(np.float32)
        obs, rew, done, info = env.step(torch.as_tensor(a, device="cuda:0"))
        rew = rew.cpu().numpy(); gstate = env.get_state(); gpart = env.get_particles(); terms = env.sim.get_reward_terms()
        for e in range(n):
            o = oracles[int(env.variants[e])]
            oobs, orew, oinfo, oc = o.step(recs[e], a[e], parts[e])
            gp = part_to_f64(gpart[e])
            for slot in (P_ALIVE, P_HIT, P_EV_EAT, P_EV_SPILL, P_EV_HIT):
                assert part_masks(gp, slot) == part_masks(parts[e], slot), (env_id, s, e, slot)
            assert np.abs(env_to_f64(gstate[e])[:32] - recs[e][:32]).max() < 1e-4
            alive = part_masks(parts[e], P_ALIVE)
            for p in range(npart):
                if (alive >> p) & 1:
                    assert max(abs(gp[P_POS + 64 * c + p] - parts[e][P_POS + 64 * c + p]) for c in range(3)) < 2e-3
            # the Drinking tilt term sits on the +-pi branch cut of the reference's roll angle at the start pose: compare the
            # reward without it unless both sides took the same branch
            if abs(terms[e, 6] - oinfo[6]) < 1e-3:
                assert abs(orew - float(rew[e])) < 2e-3
            assert abs(terms[e, 4] - oinfo[4]) < 1e-4 and abs(terms[e, 5] - oinfo[5]) < 1e-4
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", ["FeedingJaco-v0", "DrinkingJaco-v0"])
def test_gpu_particle_events_bit_exact(env_id):
    """Particles teleported to the mouth / below z = 0.5 / onto the person's lap: the eaten / spilled / hit bitmaps, the alive
    and hit masks, task_success and the food reward of the CUDA epilogue equal the oracle's (feeding.py:92-121, drinking.py:95-136)."""
    import torch
    from oracle.oracle import env_to_f64, part_to_f64, part_masks
    n = 4
    env, obs0, oracles, recs, parts = _gpu_env(env_id, n, seed=17)
    state = env.get_state(); part = env.get_particles()
    env.sim.enable_debug(True)
    a = np.zeros((n, env.sim.n_actions), dtype=np.float32)
    env.step(torch.as_tensor(a, device="cuda:0"))                                              # publishes the mouth position
    state = env.get_state(); part = env.get_particles()
    for e in range(n):
        head = oracles[int(env.variants[e])].frame(env_to_f64(state[e]), F_HEAD)
        mouth = state[e, E_TARGET_POS:E_TARGET_POS + 3] + (0.014 * quat_rot(head[3:], [0, -1, 0])).astype(np.float32)   # just in front of the lips
        for p, x, vz in ((1, mouth, 0.5396), (3, np.array([0.5, 0.5, 0.2], dtype=np.float32), 0.0),     # back at the mouth after the step / on the floor
                         (6, np.array([0.1, -0.05, 0.62 + 0.08], dtype=np.float32), 0.0)):          # above the right thigh: falls onto the lap
            for c in range(3):
                part[e, P_POS + 64 * c + p] = x[c]; part[e, P_VEL + 64 * c + p] = vz if c == 2 else 0.0
    env.sim.set_state(state, env.variants); env.sim.set_particles(part)
    recs = [env_to_f64(state[e]).copy() for e in range(n)]; parts = [part_to_f64(part[e]).copy() for e in range(n)]
    seen_hit = False
    for s in range(6):
        obs, rew, done, info = env.step(torch.as_tensor(a, device="cuda:0"))
        gstate = env.get_state(); gpart = env.get_particles(); terms = env.sim.get_reward_terms()
        for e in range(n):
            o = oracles[int(env.variants[e])]
            oobs, orew, oinfo, oc = o.step(recs[e], a[e], parts[e])
            gp = part_to_f64(gpart[e])
            for slot in (P_ALIVE, P_HIT, P_EV_EAT, P_EV_SPILL, P_EV_HIT):
                assert part_masks(gp, slot) == part_masks(parts[e], slot), (env_id, s, e, slot)
            if s == 0:
                assert part_masks(gp, P_EV_EAT) == 1 << 1 and part_masks(gp, P_EV_SPILL) == 1 << 3
            seen_hit = seen_hit or part_masks(gp, P_EV_HIT) != 0
            assert env_to_f64(gstate[e])[E_TASK_SUCCESS] == recs[e][E_TASK_SUCCESS]
            if abs(terms[e, 6] - oinfo[6]) < 1e-3:
                assert abs(orew - float(rew[e])) < 2e-3
    assert seen_hit                                                                            # the particle dropped on the lap hit the person
    env.close()


@pytest.mark.gpu
def test_gpu_every_feeding_drinking_id_steps():
    """All 16 ids (Jaco, PR2, Sawyer, Baxter x robot-only / human-active): device reset, 10 random-action steps, finite outputs,
    unit quaternions, no contact overflow, and the reference's widths (feeding.py:18)."""
    import torch
    from assistive_vr_gym_b200 import make
    for env_id in FD_IDS:
        env = make(env_id, num_envs=64, device=0, seed=3)
        obs = env.reset()
        human = env_id.endswith("Human-v0")
        assert obs.shape == (64, 48 if human else 25) and env.sim.n_actions == (11 if human else 7)
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        for _ in range(10):
            act = torch.rand((64, env.sim.n_actions), device="cuda", generator=g) * 2 - 1
            obs, rew, done, info = env.step(act)
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
        tq = obs[:, 3:7]
        assert (tq.norm(dim=1) - 1).abs().max() < 1e-4
        assert int((info["contact_overflow"] != 0).sum()) == 0, env_id
        env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("env_id,n", [("FeedingSawyer-v0", 4096), ("DrinkingBaxter-v0", 1024)])
def test_gpu_episode_invariants(env_id, n):
    """BASELINE.json configs[1] / [3] shapes: a whole 200-step random-action episode: finite state, the alive masks only lose
    bits, task_success equals the number of eaten particles, replicated environments stay bit-identical."""
    import torch
    from assistive_vr_gym_b200 import make
    env = make(env_id, num_envs=n, device=0, seed=5)
    env.reset()
    state = env.get_state(); part = env.get_particles()
    state[1] = state[0]; part[1] = part[0]                                      # environment 1 replicates environment 0
    var = env.variants.copy(); var[1] = var[0]
    env.set_state(state, var, part)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    prev_alive = env.get_particles().view(np.uint32)[:, P_ALIVE:P_ALIVE + 2].copy()
    eaten = np.zeros(n, dtype=np.int64)
    for s in range(200):
        act = torch.rand((n, env.sim.n_actions), device="cuda", generator=g) * 2 - 1
        act[1] = act[0]
        obs, rew, done, info = env.step(act)
        if s % 20 == 19 or s == 199:
            p = env.get_particles().view(np.uint32)
            alive = p[:, P_ALIVE:P_ALIVE + 2]
            assert np.all((alive & ~prev_alive) == 0)
            prev_alive = alive.copy()
        pe = env.get_particles().view(np.uint32)[:, P_EV_EAT:P_EV_EAT + 2] if s % 1 == 0 else None
        eaten += np.array([bin(int(a0)).count("1") + bin(int(a1)).count("1") for a0, a1 in pe])
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    st = env.get_state()
    assert np.isfinite(st[:, :64]).all()
    assert np.array_equal(st[:, E_TASK_SUCCESS].astype(np.int64), eaten)
    pf = env.get_particles().view(np.uint32)                                   # bit patterns: mask words are not numbers
    assert np.array_equal(st.view(np.uint32)[0], st.view(np.uint32)[1]) and np.array_equal(pf[0], pf[1])
    env.close()
