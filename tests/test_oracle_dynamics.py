"""Analytic checks of the CPU oracle's dynamics (no reference tests exist, so the oracle is pinned
against closed-form properties): symmetric positive-definite M^-1, energy conservation of the ABA + integrator with
damping, motors, contacts and the weld switched off, free-body momentum conservation, motor tracking."""
import numpy as np
import pytest

from assistive_vr_gym_b200.compiler.reset import sample_states
from assistive_vr_gym_b200.compiler.blob import read_blob
from oracle.oracle import Oracle, env_to_f64
from helpers import patch_blob, quat_rot

AVG_DOF_MOTOR, AVG_DOF_LIMIT, AVG_DOF_HARD = 2, 1, 8


@pytest.fixture(scope="module")
def states(env_data):
    env, variant = sample_states(env_data[1], 8, np.random.RandomState(5), genders=np.zeros(8, dtype=np.int32))
    return env_to_f64(env)


def test_inverse_mass_matrix_is_spd(oracles, states):
    for e in range(4):
        qdd, minv = oracles[0].dynamics(states[e].copy())
        assert np.abs(minv - minv.T).max() < 1e-9 * np.abs(minv).max()
        w = np.linalg.eigvalsh(0.5 * (minv + minv.T))
        assert w.min() > 0
        # block structure: robot (10) / human arm (7) / tool (6)
        assert np.abs(minv[:10, 10:]).max() == 0 and np.abs(minv[10:17, 17:]).max() == 0


def test_static_equilibrium_without_gravity(env_data, states):
    """At rest with zero gravity everywhere the unconstrained acceleration vanishes."""
    blob = patch_blob(env_data[0][0], zero_gravity=True)
    o = Oracle(blob)
    qdd, _ = o.dynamics(states[0].copy())
    assert np.abs(qdd).max() < 1e-12


def _energy(o, model, rec):
    qdd, minv = o.dynamics(rec)
    M = np.linalg.inv(minv)
    qd = rec[32:32 + o.n_dof]
    ke = 0.5 * qd @ M @ qd
    pe = 0.0
    for b, body in enumerate(model["bodies"]):
        g = np.asarray(body["gravity"], dtype=np.float64)
        pe -= float(body["mass"]) * g @ o.body_pose(rec, b)[:3]
    return ke + pe


def test_energy_conservation_free_swing(env_data, states):
    """Human arm swinging under its gravity, robot coasting, tool spinning: with damping, motors, limits, contacts and
    the weld removed, E = 1/2 qd^T M qd + potential is conserved up to the O(dt) error of semi-implicit Euler."""
    blob = patch_blob(env_data[0][0], header={"lin_damp": 0.0, "ang_damp": 0.0, "n_pair": 0, "weld_max_force": 0.0, "dt": 0.0005,
                                               "substeps": 1, "solver_iters": 1},
                      dof_flags_clear=AVG_DOF_MOTOR | AVG_DOF_LIMIT | AVG_DOF_HARD)
    o = Oracle(blob)
    model = read_blob(blob)
    rec = states[1].copy()
    rng = np.random.RandomState(0)
    rec[32:32 + 17] = rng.uniform(-0.5, 0.5, 17)           # joint velocities
    rec[32 + 17:32 + 23] = rng.uniform(-0.3, 0.3, 6)       # tool twist
    e0 = _energy(o, model, rec)
    act = np.zeros(7, dtype=np.float32)
    es = []
    for _ in range(400):
        o.step(rec, act)
        es.append(_energy(o, model, rec))
    drift = np.abs(np.array(es) - e0).max() / abs(e0)
    assert drift < 5e-3, drift


def test_free_body_momentum(env_data, states):
    """Torque-free tool: linear velocity constant, world angular momentum R I R^T w conserved."""
    blob = patch_blob(env_data[0][0], header={"lin_damp": 0.0, "ang_damp": 0.0, "n_pair": 0, "weld_max_force": 0.0, "dt": 0.001,
                                               "substeps": 1}, dof_flags_clear=0)
    o = Oracle(blob)
    model = read_blob(blob)
    tb = [i for i, b in enumerate(model["bodies"]) if b["jtype"] == 2][0]
    body = model["bodies"][tb]
    rec = states[2].copy()
    rec[32 + 17:32 + 23] = [0.1, -0.2, 0.05, 3.0, -2.0, 1.0]

    def ang_mom(rec):
        q = rec[int(body["qidx"]) + 3:int(body["qidx"]) + 7]
        w = rec[32 + 20:32 + 23]
        wl = quat_rot([-q[0], -q[1], -q[2], q[3]], w)
        return quat_rot(q, np.asarray(body["inertia"], dtype=np.float64) * wl)

    l0 = ang_mom(rec); v0 = rec[32 + 17:32 + 20].copy()
    for _ in range(300):
        o.step(rec, np.zeros(7, dtype=np.float32))
    assert np.abs(rec[32 + 17:32 + 20] - v0).max() < 1e-12
    assert np.abs(ang_mom(rec) - l0).max() < 2e-2 * np.linalg.norm(l0)


def test_position_motor_reaches_target_velocity(env_data, states):
    """A position motor row drives its joint to kp*(q*-q)/dt in one solve when unclamped (kd = 1)."""
    blob = patch_blob(env_data[0][0], header={"n_pair": 0, "substeps": 1, "residual_thr": 0.0, "solver_iters": 200, "weld_max_force": 0.0},
                      zero_gravity=True)
    o = Oracle(blob)
    rec = states[3].copy()
    q0 = rec[:7].copy()
    act = np.array([0.1, -0.1, 0.2, 0, 0, 0, 0], dtype=np.float32)     # small enough that the 1 N m clamp is inactive
    o.step(rec, act)
    target = q0 + np.clip(act, -1, 1).astype(np.float32) * np.float32(0.05) * 1      # frame_skip patched to 1
    expect = 0.05 * (target - q0) / 0.02
    assert np.abs(rec[64:71] - target).max() < 1e-6
    assert np.abs(rec[32:39] - expect).max() < 2e-3
