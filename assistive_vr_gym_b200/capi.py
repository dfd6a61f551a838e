"""ctypes binding of the C-ABI in include/avg_b200.h (libavg_b200.so).

This is the same stub a maintainer of the reference would add in place of `import pybullet as p` on the step path
(INTEGRATION.md).  It fails loudly when the CUDA library is missing or no GPU is present: there is no CPU path.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AVG_B200_LIB") or os.path.join(_HERE, "libavg_b200.so")   # override: A/B builds during tuning

EXPORTS = [
    "avg_create", "avg_destroy", "avg_last_error", "avg_upload_model", "avg_set_state", "avg_get_state", "avg_get_variants", "avg_set_time_limit",
    "avg_state_device_ptr", "avg_reset_obs", "avg_step", "avg_step_host", "avg_enable_debug", "avg_get_contacts",
    "avg_get_reward_terms", "avg_num_envs", "avg_num_actions", "avg_num_obs", "avg_env_stride", "avg_launch_count",
    "avg_bytes_per_env_step", "avg_arm_limit_logits", "avg_alloc_host", "avg_free_host", "avg_upload_reset_table", "avg_reset",
    "avg_upload_policy", "avg_policy_act",
    "avg_set_particles", "avg_get_particles", "avg_particles_device_ptr", "avg_particle_stride", "avg_num_particles", "avg_settle",
]

CONTACT_DT = np.dtype([("shape_a", "<i4"), ("shape_b", "<i4"), ("pos_a", "<f4", 3), ("pos_b", "<f4", 3),
                       ("normal", "<f4", 3), ("dist", "<f4"), ("force", "<f4"), ("pad", "<i4", 3)])
assert CONTACT_DT.itemsize == 64
MAX_CONTACT = 32       # AVG_MAX_CONTACT (include/avg_model.h)


class AvgError(RuntimeError):
    pass


_lib = None


def load_library(build_if_missing: bool = True) -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise AvgError(f"{LIB_PATH} is missing; run `python -c 'import __graft_entry__ as g; g.build()'`")
        from .build import build
        build()
    lib = ctypes.CDLL(LIB_PATH)
    vp, ip, fp = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_float)
    lib.avg_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]
    lib.avg_destroy.argtypes = [vp]
    lib.avg_last_error.argtypes = [vp]; lib.avg_last_error.restype = ctypes.c_char_p
    lib.avg_upload_model.argtypes = [vp, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
    lib.avg_set_state.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp]
    lib.avg_get_state.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp]
    lib.avg_get_variants.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp]
    lib.avg_set_time_limit.argtypes = [vp, ctypes.c_int]
    lib.avg_state_device_ptr.argtypes = [vp]; lib.avg_state_device_ptr.restype = vp
    lib.avg_reset_obs.argtypes = [vp, vp, vp]
    lib.avg_step.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.avg_step_host.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.avg_enable_debug.argtypes = [vp, ctypes.c_int]
    lib.avg_get_contacts.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp]
    lib.avg_get_reward_terms.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp]
    lib.avg_upload_reset_table.argtypes = [vp, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
    lib.avg_reset.argtypes = [vp, vp, ctypes.c_uint32, vp, vp]
    lib.avg_upload_policy.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t]
    lib.avg_policy_act.argtypes = [vp, vp, vp, vp]
    lib.avg_alloc_host.argtypes = [ctypes.c_size_t, ctypes.POINTER(vp)]
    lib.avg_free_host.argtypes = [vp]
    lib.avg_arm_limit_logits.argtypes = [vp, ctypes.c_int, vp, vp, ctypes.c_int, vp]
    lib.avg_set_particles.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp]
    lib.avg_get_particles.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp]
    lib.avg_particles_device_ptr.argtypes = [vp]; lib.avg_particles_device_ptr.restype = vp
    lib.avg_particle_stride.argtypes = []
    lib.avg_num_particles.argtypes = [vp]
    lib.avg_settle.argtypes = [vp, vp, ctypes.c_int, vp]
    for f in ("avg_num_envs", "avg_num_actions", "avg_num_obs", "avg_bytes_per_env_step"):
        getattr(lib, f).argtypes = [vp]
    lib.avg_env_stride.argtypes = []
    lib.avg_launch_count.argtypes = [vp]; lib.avg_launch_count.restype = ctypes.c_longlong
    _lib = lib
    return lib


class PinnedArray:
    """NumPy view of page-locked host memory from avg_alloc_host (freed with the object)."""

    def __init__(self, shape, dtype):
        self.lib = load_library()
        dt = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dt.itemsize
        self.ptr = ctypes.c_void_p()
        if self.lib.avg_alloc_host(max(nbytes, 1), ctypes.byref(self.ptr)) != 0:
            raise AvgError("avg_alloc_host failed")
        buf = (ctypes.c_char * max(nbytes, 1)).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
        self.array[...] = 0

    def __del__(self):
        try:
            if self.ptr:
                self.array = None
                self.lib.avg_free_host(self.ptr); self.ptr = None
        except Exception:
            pass


class Sim:
    """Thin object wrapper of one AvgHandle (one GPU)."""

    def __init__(self, n_env: int, device: int = 0):
        self.lib = load_library()
        self.h = ctypes.c_void_p()
        rc = self.lib.avg_create(device, n_env, ctypes.byref(self.h))
        if rc != 0:
            raise AvgError(f"avg_create failed ({rc}): {self.lib.avg_last_error(None).decode()}")
        self.n_env = n_env
        self.device = device

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise AvgError(f"{what} failed ({rc}): {self.lib.avg_last_error(self.h).decode()}")

    def close(self):
        if self.h:
            self.lib.avg_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_model(self, variant: int, blob: bytes):
        self._check(self.lib.avg_upload_model(self.h, variant, blob, len(blob)), "avg_upload_model")

    def set_state(self, env: np.ndarray, variants: np.ndarray | None = None, begin: int = 0):
        env = np.ascontiguousarray(env, dtype=np.float32)
        v = None if variants is None else np.ascontiguousarray(variants, dtype=np.int32)
        self._check(self.lib.avg_set_state(self.h, begin, env.shape[0], env.ctypes.data, None if v is None else v.ctypes.data),
                    "avg_set_state")

    def get_state(self, begin: int = 0, count: int | None = None) -> np.ndarray:
        count = self.n_env - begin if count is None else count
        out = np.zeros((count, self.lib.avg_env_stride()), dtype=np.float32)
        self._check(self.lib.avg_get_state(self.h, begin, count, out.ctypes.data), "avg_get_state")
        return out

    def set_time_limit(self, max_episode_steps: int):
        self._check(self.lib.avg_set_time_limit(self.h, int(max_episode_steps)), "avg_set_time_limit")

    def set_particles(self, part: np.ndarray, begin: int = 0):
        part = np.ascontiguousarray(part, dtype=np.float32)
        assert part.shape[1] == self.lib.avg_particle_stride()
        self._check(self.lib.avg_set_particles(self.h, begin, part.shape[0], part.ctypes.data), "avg_set_particles")

    def get_particles(self, begin: int = 0, count: int | None = None) -> np.ndarray:
        count = self.n_env - begin if count is None else count
        out = np.zeros((count, self.lib.avg_particle_stride()), dtype=np.float32)
        self._check(self.lib.avg_get_particles(self.h, begin, count, out.ctypes.data), "avg_get_particles")
        return out

    def settle(self, n_steps: int, mask_ptr: int = 0, stream: int = 0):
        """n_steps x p.stepSimulation() without actions (the settle loop of reset(), feeding.py:318-320)."""
        self._check(self.lib.avg_settle(self.h, mask_ptr or None, int(n_steps), stream), "avg_settle")

    @property
    def n_particles(self) -> int:
        return self.lib.avg_num_particles(self.h)

    @property
    def state_ptr(self) -> int:
        return int(self.lib.avg_state_device_ptr(self.h) or 0)

    def get_state_tensor(self, torch, device):
        """Zero-copy [n_env, AVG_ENV_STRIDE] float32 CUDA tensor over the state arena (avg_state_device_ptr)."""
        if getattr(self, "_state_tensor", None) is None:
            stride = self.lib.avg_env_stride()

            class _Arena:
                __cuda_array_interface__ = {"shape": (self.n_env, stride), "typestr": "<f4", "data": (self.state_ptr, False), "version": 2}
            self._state_tensor = torch.as_tensor(_Arena(), device=device)
        return self._state_tensor

    def get_variants(self, begin: int = 0, count: int | None = None) -> np.ndarray:
        count = self.n_env - begin if count is None else count
        out = np.zeros(count, dtype=np.int32)
        self._check(self.lib.avg_get_variants(self.h, begin, count, out.ctypes.data), "avg_get_variants")
        return out

    def upload_reset_table(self, variant: int, table: bytes):
        self._check(self.lib.avg_upload_reset_table(self.h, variant, table, len(table)), "avg_upload_reset_table")

    def reset_device(self, mask_ptr: int, seed: int, obs_ptr: int, stream: int = 0):
        """avg_reset: mask_ptr = device uint8[n_env] or 0 (all)."""
        self._check(self.lib.avg_reset(self.h, mask_ptr or None, seed & 0xffffffff, obs_ptr or None, stream), "avg_reset")

    def upload_policy(self, blob: bytes):
        self._check(self.lib.avg_upload_policy(self.h, blob, len(blob)), "avg_upload_policy")

    def policy_act(self, obs_ptr: int, act_ptr: int, stream: int = 0):
        self._check(self.lib.avg_policy_act(self.h, obs_ptr, act_ptr, stream), "avg_policy_act")

    def reset_obs(self, obs_ptr: int, stream: int = 0):
        self._check(self.lib.avg_reset_obs(self.h, obs_ptr, stream), "avg_reset_obs")

    def step(self, act_ptr: int, obs_ptr: int, rew_ptr: int, done_ptr: int, info_ptr: int, stream: int = 0):
        self._check(self.lib.avg_step(self.h, act_ptr, obs_ptr, rew_ptr, done_ptr, info_ptr, stream), "avg_step")

    def step_host(self, actions: np.ndarray, obs: np.ndarray, reward: np.ndarray, done: np.ndarray, info: np.ndarray):
        self._check(self.lib.avg_step_host(self.h, actions.ctypes.data, obs.ctypes.data, reward.ctypes.data,
                                           done.ctypes.data, info.ctypes.data), "avg_step_host")

    def enable_debug(self, on: bool = True):
        self._check(self.lib.avg_enable_debug(self.h, int(on)), "avg_enable_debug")

    def get_contacts(self, begin: int = 0, count: int | None = None):
        count = self.n_env - begin if count is None else count
        c = np.zeros((count, MAX_CONTACT), dtype=CONTACT_DT)
        n = np.zeros(count, dtype=np.int32)
        self._check(self.lib.avg_get_contacts(self.h, begin, count, c.ctypes.data, n.ctypes.data), "avg_get_contacts")
        return c, n

    def get_reward_terms(self, begin: int = 0, count: int | None = None) -> np.ndarray:
        count = self.n_env - begin if count is None else count
        t = np.zeros((count, 8), dtype=np.float32)
        self._check(self.lib.avg_get_reward_terms(self.h, begin, count, t.ctypes.data), "avg_get_reward_terms")
        return t

    def arm_limit_logits(self, variant: int, q4_ptr: int, out_ptr: int, n: int, stream: int = 0):
        self._check(self.lib.avg_arm_limit_logits(self.h, variant, q4_ptr, out_ptr, n, stream), "avg_arm_limit_logits")

    @property
    def n_actions(self) -> int:
        return self.lib.avg_num_actions(self.h)

    @property
    def n_obs(self) -> int:
        return self.lib.avg_num_obs(self.h)

    @property
    def launch_count(self) -> int:
        return int(self.lib.avg_launch_count(self.h))

    @property
    def bytes_per_env_step(self) -> int:
        return int(self.lib.avg_bytes_per_env_step(self.h))
