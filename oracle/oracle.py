"""ctypes wrapper of the CPU oracle (oracle/avg_oracle.c).

TEST INFRASTRUCTURE ONLY — may be imported from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs, never from the product package. PARITY UNPINNED (see the header of avg_oracle.c).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libavg_oracle.so")
ENV_STRIDE = 192
INT_SLOTS = (123, 152, 161, 166, 167, 168)    # AVG_E_LIMB_FRAME, AVG_E_ITERATION, AVG_E_HAS_VALID, AVG_E_OVERFLOW
UINT_SLOTS = (170, 171, 172, 173, 174, 175,   # AVG_E_TARGET_MASK words (BedBathing), AVG_E_FROZEN
              176, 178, 180, 182, 184, 186, 188, 190)   # keys of the warm-start contact cache (AVG_E_WCACHE)
P_STRIDE = 592                                # AVG_P_STRIDE (include/avg_model.h)
P_UINT_SLOTS = tuple(range(576, 590))         # AVG_P_ALIVE .. AVG_P_EV_HIT mask words
P_INT_SLOTS = (590, 591)                      # AVG_P_NCONTACT, overflow flags
MAX_CONTACT = 32                              # AVG_MAX_CONTACT
_DP = ctypes.POINTER(ctypes.c_double)
_FP = ctypes.POINTER(ctypes.c_float)
_IP = ctypes.POINTER(ctypes.c_int)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "avg_oracle.c"), os.path.join(_HERE, "..", "include", "avg_model.h")]
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(os.path.getmtime(p) for p in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


def env_to_f64(env_f32: np.ndarray) -> np.ndarray:
    """float32 device record (ints bit-cast) -> float64 oracle record (ints as numbers)."""
    env_f32 = np.ascontiguousarray(env_f32, dtype=np.float32)
    out = env_f32.astype(np.float64)
    iv = env_f32.view(np.int32)
    for s in INT_SLOTS:
        out[..., s] = iv[..., s]
    uv = env_f32.view(np.uint32)
    for s in UINT_SLOTS:
        out[..., s] = uv[..., s]
    return out


def env_to_f32(env_f64: np.ndarray) -> np.ndarray:
    out = env_f64.astype(np.float32)
    iv = out.view(np.int32)
    for s in INT_SLOTS:
        iv[..., s] = np.rint(env_f64[..., s]).astype(np.int32)
    uv = out.view(np.uint32)
    for s in UINT_SLOTS:
        uv[..., s] = np.rint(env_f64[..., s]).astype(np.uint32)
    return out


def part_to_f64(part_f32: np.ndarray) -> np.ndarray:
    """float32 device particle record (masks bit-cast) -> float64 oracle record (masks as numbers)."""
    part_f32 = np.ascontiguousarray(part_f32, dtype=np.float32)
    out = part_f32.astype(np.float64)
    uv = part_f32.view(np.uint32); iv = part_f32.view(np.int32)
    for s in P_UINT_SLOTS:
        out[..., s] = uv[..., s]
    for s in P_INT_SLOTS:
        out[..., s] = iv[..., s]
    return out


def part_masks(part_f64: np.ndarray, slot: int) -> int:
    """64-bit mask stored in two words at `slot` of a float64 particle record."""
    return int(part_f64[slot]) | (int(part_f64[slot + 1]) << 32)


class Oracle:
    def __init__(self, blob: bytes):
        self.lib = ctypes.CDLL(build())
        self.blob = ctypes.create_string_buffer(blob, len(blob))
        from assistive_vr_gym_b200.compiler.blob import read_blob
        self.model = read_blob(blob)
        h = self.model["header"]
        self.n_obs = int(h["n_obs_robot"] + h["n_obs_human"])
        self.n_act = int(h["n_action_robot"] + h["n_action_human"])
        self.n_dof = int(h["n_dof"])
        L = self.lib
        L.avg_oracle_step.restype = ctypes.c_int
        L.avg_oracle_step.argtypes = [ctypes.c_void_p, _DP, _FP, _DP, _DP, _DP, _DP, _IP]
        L.avg_oracle_reset_obs.argtypes = [ctypes.c_void_p, _DP, _DP]
        L.avg_oracle_frame.argtypes = [ctypes.c_void_p, _DP, ctypes.c_int, _DP]
        L.avg_oracle_body_pose.argtypes = [ctypes.c_void_p, _DP, ctypes.c_int, _DP]
        L.avg_oracle_dynamics.argtypes = [ctypes.c_void_p, _DP, _DP, _DP]
        L.avg_oracle_collide.argtypes = [ctypes.c_void_p, _DP, _DP, _IP]
        L.avg_oracle_shape_pair.argtypes = [ctypes.c_void_p, ctypes.c_int, _DP, ctypes.c_int, _DP, ctypes.c_double, _DP]
        L.avg_oracle_sizes.argtypes = [_IP]
        L.avg_oracle_arm_limit.argtypes = [ctypes.c_void_p, _DP, _DP]
        L.avg_oracle_step_fd.restype = ctypes.c_int
        L.avg_oracle_step_fd.argtypes = [ctypes.c_void_p, _DP, _DP, _FP, _DP, _DP, _DP, _DP, _IP]
        L.avg_oracle_settle.argtypes = [ctypes.c_void_p, _DP, _DP, ctypes.c_int]
        L.avg_oracle_particle_collide.argtypes = [ctypes.c_void_p, _DP, _DP, _DP, _IP]
        self.n_particle = int(h["n_particle"])

    def sizes(self):
        a = (ctypes.c_int * 8)()
        n = self.lib.avg_oracle_sizes(a)
        return list(a)[:n]

    @staticmethod
    def _dp(a):
        return a.ctypes.data_as(_DP)

    def reset_obs(self, env: np.ndarray) -> np.ndarray:
        obs = np.zeros(self.n_obs)
        assert self.lib.avg_oracle_reset_obs(self.blob, self._dp(env), self._dp(obs)) == 0
        return obs

    def step(self, env: np.ndarray, action: np.ndarray, part: np.ndarray | None = None):
        """env: float64 [ENV_STRIDE] (modified in place); part: float64 [P_STRIDE] particle record of Feeding / Drinking
        (modified in place). -> obs, reward, info[8], contacts[n,13]"""
        assert env.dtype == np.float64 and env.flags.c_contiguous
        act = np.ascontiguousarray(action, dtype=np.float32)
        obs = np.zeros(self.n_obs)
        rew = np.zeros(1)
        info = np.zeros(8)
        cont = np.zeros((MAX_CONTACT, 13))
        nc = ctypes.c_int(0)
        if part is not None:
            assert part.dtype == np.float64 and part.flags.c_contiguous and part.shape == (P_STRIDE,)
        rc = self.lib.avg_oracle_step_fd(self.blob, self._dp(env), self._dp(part) if part is not None else None,
                                         act.ctypes.data_as(_FP), self._dp(obs), self._dp(rew),
                                         self._dp(info), self._dp(cont), ctypes.byref(nc))
        assert rc == 0, rc
        return obs, float(rew[0]), info, cont[:nc.value].copy()

    def settle(self, env: np.ndarray, part: np.ndarray | None, n: int):
        """n x p.stepSimulation() without actions (the reset() settle loop, feeding.py:318-320)."""
        assert self.lib.avg_oracle_settle(self.blob, self._dp(env), self._dp(part) if part is not None else None, int(n)) == 0

    def particle_collide(self, env: np.ndarray, part: np.ndarray):
        out = np.zeros((320, 8)); n = ctypes.c_int(0)
        self.lib.avg_oracle_particle_collide(self.blob, self._dp(env), self._dp(part), self._dp(out), ctypes.byref(n))
        return out[:n.value].copy()

    def frame(self, env, f):
        out = np.zeros(7)
        self.lib.avg_oracle_frame(self.blob, self._dp(env), f, self._dp(out))
        return out

    def body_pose(self, env, b):
        out = np.zeros(7)
        self.lib.avg_oracle_body_pose(self.blob, self._dp(env), b, self._dp(out))
        return out

    def dynamics(self, env):
        qdd = np.zeros(32)
        minv = np.zeros((self.n_dof, self.n_dof))
        self.lib.avg_oracle_dynamics(self.blob, self._dp(env), self._dp(qdd), self._dp(minv))
        return qdd[:self.n_dof], minv

    def collide(self, env):
        cont = np.zeros((MAX_CONTACT, 13))
        nc = ctypes.c_int(0)
        self.lib.avg_oracle_collide(self.blob, self._dp(env), self._dp(cont), ctypes.byref(nc))
        return cont[:nc.value].copy()

    def arm_limit_logit(self, q4):
        q = np.ascontiguousarray(q4, dtype=np.float64); out = np.zeros(1)
        rc = self.lib.avg_oracle_arm_limit(self.blob, self._dp(q), self._dp(out))
        assert rc == 0, rc
        return float(out[0])

    def shape_pair(self, sa, pose_a, sb, pose_b, thr=1e9):
        out = np.zeros(10)
        pa = np.ascontiguousarray(pose_a, dtype=np.float64); pb = np.ascontiguousarray(pose_b, dtype=np.float64)
        hit = self.lib.avg_oracle_shape_pair(self.blob, sa, self._dp(pa), sb, self._dp(pb), float(thr), self._dp(out))
        return hit, out
