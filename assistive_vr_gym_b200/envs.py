"""Host-side mirror of the reference's gym interface, batched over N environments resident on one GPU.

Keeps the `assistive_gym` surface: ids `'<Task><Robot>[Human]-v0'` (reference
`assistive_gym/__init__.py:4-344`), `seed`, `reset() -> obs`, `step(action) -> (obs, reward, done, info)` with the
reference's observation layout (`scratch_itch.py:104-128`), reward (`:62-72`) and info keys (`:77`), and the
`TimeLimit(200)` wrapper `gym.make` adds (`__init__.py:18-22`).  All arithmetic happens in the CUDA library behind
`include/avg_b200.h`; PyTorch only owns the I/O tensors and the stream.
"""
from __future__ import annotations

import os
import re
from typing import Dict, List, Optional

import numpy as np

from . import capi
from .compiler.reset import device_ik_setup, reset_table_bytes, sample_states
from .compiler.reset_fd import sample_states_fd

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
MAX_EPISODE_STEPS = 200      # reference __init__.py:21

# ids registered by the reference for the tasks/robots this round compiles (reference __init__.py:16-50)
REGISTRY: Dict[str, dict] = {
    "ScratchItchJaco-v0": dict(task="scratch_itch", robot="jaco", human_control=False, data="ScratchItchJaco.npz"),
    "ScratchItchJacoHuman-v0": dict(task="scratch_itch", robot="jaco", human_control=True, data="ScratchItchJacoHuman.npz"),
    # reference __init__.py:103-108; observation 24 wide (bed_bathing.py:19,147)
    "BedBathingJaco-v0": dict(task="bed_bathing", robot="jaco", human_control=False, data="BedBathingJaco.npz"),
    "BedBathingJacoHuman-v0": dict(task="bed_bathing", robot="jaco", human_control=True, data="BedBathingJacoHuman.npz"),
    # PR2 ids (reference __init__.py:4-14, 91-101): left arm + gripper integrated, the other branches baked static per base pose
    "ScratchItchPR2-v0": dict(task="scratch_itch", robot="pr2", human_control=False, data="ScratchItchPR2.npz"),
    "ScratchItchPR2Human-v0": dict(task="scratch_itch", robot="pr2", human_control=True, data="ScratchItchPR2Human.npz"),
    "BedBathingPR2-v0": dict(task="bed_bathing", robot="pr2", human_control=False, data="BedBathingPR2.npz"),
    "BedBathingPR2Human-v0": dict(task="bed_bathing", robot="pr2", human_control=True, data="BedBathingPR2Human.npz"),
}
# Feeding / Drinking (reference __init__.py:175-260, 259-344): Jaco and PR2 are the reference's ids; the Sawyer and Baxter ids
# are named by BASELINE.json but have no reference environment: their recipes are defined by
# compiler/scene_fd.py ROBOT_FD.
for _t, _task in (("Feeding", "feeding"), ("Drinking", "drinking")):
    for _r, _robot in (("Jaco", "jaco"), ("PR2", "pr2"), ("Sawyer", "sawyer"), ("Baxter", "baxter")):
        REGISTRY[f"{_t}{_r}-v0"] = dict(task=_task, robot=_robot, human_control=False, data=f"{_t}{_r}.npz")
        REGISTRY[f"{_t}{_r}Human-v0"] = dict(task=_task, robot=_robot, human_control=True, data=f"{_t}{_r}Human.npz")
# `New` ids (reference __init__.py:38-50): a person of random height with a random waist pose (both per model variant here: 16
# per gender), no impairment, the arm pose drawn per episode on the device until it is collision-free (avg_reset_new_kernel)
REGISTRY["ScratchItchJacoNew-v0"] = dict(task="scratch_itch", robot="jaco", human_control=False, data="ScratchItchJacoNew.npz", new=True)
REGISTRY["ScratchItchPR2New-v0"] = dict(task="scratch_itch", robot="pr2", human_control=False, data="ScratchItchPR2New.npz", new=True)
REGISTRY["BedBathingJacoNew-v0"] = dict(task="bed_bathing", robot="jaco", human_control=False, data="BedBathingJacoNew.npz", new=True)
REGISTRY["BedBathingPR2New-v0"] = dict(task="bed_bathing", robot="pr2", human_control=False, data="BedBathingPR2New.npz", new=True)
# Feeding / Drinking `New` ids (__init__.py:206-218,290-302): the whole person is static in these (feeding.py:235), so height and
# waist pose per variant are all there is to them besides human_impairment = 'none'
for _t, _task in (("Feeding", "feeding"), ("Drinking", "drinking")):
    for _r, _robot in (("Jaco", "jaco"), ("PR2", "pr2")):
        REGISTRY[f"{_t}{_r}New-v0"] = dict(task=_task, robot=_robot, human_control=False, data=f"{_t}{_r}New.npz", new=True)
_OBS_LEN = {"scratch_itch": (30, 34), "bed_bathing": (24, 28), "feeding": (25, 23), "drinking": (25, 23)}      # (robot, human) widths: scratch_itch.py:19, bed_bathing.py:19, feeding.py:18
_ACT_HUMAN = {"scratch_itch": 10, "bed_bathing": 10, "feeding": 4, "drinking": 4}                              # scratch_itch.py:19, feeding.py:18
SETTLE_STEPS = 100           # feeding.py:318-320
# every non-VR id of the reference is registered above; its VR ids (`<Task>VR<Robot>[Human|New]-v0`, __init__.py:52-89 etc.) drive a
# headset and are out of scope (north_star)
_VR_REFERENCE_IDS = [f"{t}VR{r}{v}-v0" for t in ("ScratchItch", "BedBathing", "Feeding", "Drinking")
                     for r in ("PR2", "Jaco") for v in ("", "Human", "New")]


class Box:
    """Minimal stand-in for gym.spaces.Box (gym is not installed here): Box(-1, 1, (n,), float32), env.py:34-35."""

    def __init__(self, n: int, seed: int = 0):
        self.low = -np.ones(n, dtype=np.float32)
        self.high = np.ones(n, dtype=np.float32)
        self.shape = (n,)
        self.dtype = np.float32
        self._rng = np.random.RandomState(seed)

    def sample(self) -> np.ndarray:
        return self._rng.uniform(self.low, self.high).astype(np.float32)

    def seed(self, s):
        self._rng = np.random.RandomState(s)


def load_env_data(name: str):
    path = os.path.join(_DATA, name)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run tools/compile_models.py where the reference assets are available")
    z = np.load(path)
    blobs, resets = [], []
    v = 0
    while f"blob_{v}" in z:
        blobs.append(z[f"blob_{v}"].tobytes())
        pre = f"reset_{v}_"
        resets.append({k[len(pre):]: z[k] for k in z.files if k.startswith(pre)})
        v += 1
    return blobs, resets


class BatchedAssistiveEnv:
    """N copies of one reference environment stepping in lock-step on one GPU."""

    def __init__(self, env_id: str, num_envs: int = 1, device: int = 0, seed: int = 1001, auto_reset: bool = False,
                 device_ik: bool = True, cuda_graph: bool = False):
        if env_id not in REGISTRY:
            if env_id in _VR_REFERENCE_IDS:
                raise NotImplementedError(f"{env_id}: the reference's VR / headset ids are out of scope of this simulator "
                                          f"(built: {sorted(REGISTRY)})")
            raise KeyError(f"unknown environment id {env_id}")
        import torch
        if not torch.cuda.is_available():
            raise capi.AvgError("no CUDA device: the simulator has no CPU path")
        self.torch = torch
        self.spec = dict(REGISTRY[env_id]); self.env_id = env_id
        self.num_envs = int(num_envs)
        self.device_index = device
        self.device = torch.device("cuda", device)
        self.blobs, self.reset_data = load_env_data(self.spec["data"])
        self.sim = capi.Sim(self.num_envs, device)
        # device_ik (default): `reset_device` / `reset()` solve every episode's start pose on the GPU for a freshly drawn start
        # target (scratch_itch.py:243-253), as the reference does; False picks one of the pool's <= 64 precomputed IK solutions.
        # BedBathing's start target is fixed (bed_bathing.py:315), so its pool entry already is the solution and the option
        # changes nothing there; Feeding / Drinking always solve on the device.
        self.device_ik = bool(device_ik)
        for v, b in enumerate(self.blobs):
            self.sim.upload_model(v, b)
            ik = device_ik_setup(b, self.spec["task"], self.spec["robot"]) if self.device_ik else None
            self.sim.upload_reset_table(v, reset_table_bytes(self.reset_data[v], ik))
        self.action_robot_len = 7
        self.action_human_len = _ACT_HUMAN[self.spec["task"]] if self.spec["human_control"] else 0
        self.has_particles = self.spec["task"] in ("feeding", "drinking")
        self.obs_robot_len = _OBS_LEN[self.spec["task"]][0]
        self.obs_human_len = _OBS_LEN[self.spec["task"]][1] if self.spec["human_control"] else 0
        assert self.sim.n_actions == self.action_robot_len + self.action_human_len
        assert self.sim.n_obs == self.obs_robot_len + self.obs_human_len
        self.action_space = Box(self.sim.n_actions)
        self.observation_space = Box(self.sim.n_obs)
        # auto_reset: False | True (lock-stepped episodes: the whole batch restarts when the shared TimeLimit counter hits 200)
        # | "device" (staggered episodes: TimeLimit per environment on the device, the environments whose `done` byte is
        # set restart inside step() through avg_reset with that byte array as the mask -- no host round trip)
        self.auto_reset = auto_reset
        if auto_reset == "device":
            self.sim.set_time_limit(MAX_EPISODE_STEPS)
        self.seed(seed)
        n = self.num_envs
        with torch.cuda.device(self.device):
            self.obs = torch.zeros((n, self.sim.n_obs), dtype=torch.float32, device=self.device)
            self.reward = torch.zeros(n, dtype=torch.float32, device=self.device)
            self.done_dev = torch.zeros(n, dtype=torch.uint8, device=self.device)
            self.info_dev = torch.zeros((n, 2), dtype=torch.float32, device=self.device)
        self.elapsed = 0
        self.variants: Optional[np.ndarray] = None
        self._needs_reset = True
        # cuda_graph: the launch sequence of one step (44 kernels at two half batches, plus the device auto-reset and, in
        # rollout(), the policy kernel) is captured once into a CUDA graph and replayed -- one driver call per step instead
        # of one per kernel.  Pays where the step is launch-bound (small batches); results are bit-identical to eager steps.
        self.cuda_graph = bool(cuda_graph)
        self._graphs = {}
        self._eager_steps = 0

    # -- CUDA graph of the step ---------------------------------------------------------------------------------------
    def _enqueue_step(self, act_ptr: int):
        """The device work of one step on the current stream: avg_step and, with auto_reset="device", the masked restart."""
        torch = self.torch
        self.sim.step(act_ptr, self.obs.data_ptr(), self.reward.data_ptr(), self.done_dev.data_ptr(),
                      self.info_dev.data_ptr(), self._stream())
        if self.auto_reset == "device":
            self.terminal_obs.copy_(self.obs)                       # last observation of the episodes that end at this step
            self.sim.reset_device(self.done_dev.data_ptr(), self._device_seed, self.obs.data_ptr(), self._stream())

    def _replay(self, key: str, with_policy: bool):
        """Replay (capturing on first use) the graph `key`: [policy ->] step, reading the static action buffer."""
        torch = self.torch
        g = self._graphs.get(key)
        if g is None:
            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(cur)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                side.synchronize()
                # thread-local capture mode: other threads (NCCL's watchdog under torchrun) keep making CUDA calls meanwhile
                g.capture_begin(capture_error_mode="thread_local")
                try:
                    if with_policy:
                        self.sim.policy_act(self.obs.data_ptr(), self.actions_dev.data_ptr(), self._stream())
                    self._enqueue_step(self.actions_dev.data_ptr())
                finally:
                    g.capture_end()
            cur.wait_stream(side)
            self._graphs[key] = g
        g.replay()

    # -- reference API ------------------------------------------------------------------------------------------
    def seed(self, seed=None):
        """env.py:80-82"""
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def reset(self, genders: Optional[np.ndarray] = None, host: Optional[bool] = None):
        """scratch_itch.py:130-273, batched.  Default: the device sampler (`reset_device`: every draw of the reference's reset,
        a fresh start target per episode with the IK solved on the GPU).  host=True (implied by explicit `genders`): the numpy
        sampler of compiler/reset.py, whose start poses come from the variant's pool of precomputed IK solutions."""
        if host is None:
            host = genders is not None
        if host and self.spec.get("new"):
            raise NotImplementedError(f"{self.env_id}: the `New` ids draw their collision-free arm pose on the device; use reset() / reset_device()")
        if not host:
            obs = self.reset_device()
            self.variants = self.sim.get_variants()
            return obs
        if self.has_particles:
            # Feeding / Drinking: a complete draw (bowl, start target, arm pose) from the host pool, the particle grid above the
            # tool, then the reference's 100 settle steps on the device (feeding.py:318-320)
            env, part, variant = sample_states_fd(self.reset_data, self.num_envs, self.np_random, genders)
            self.set_state(env, variant, part, settle=SETTLE_STEPS)
        else:
            env, variant = sample_states(self.reset_data, self.num_envs, self.np_random, genders)
            self.set_state(env, variant)
        self._last_reset = "host"
        return self.obs

    def reset_device(self, mask=None, seed: Optional[int] = None):
        """Reset on the GPU (avg_reset): `mask` = CUDA bool/uint8 tensor [N] of environments to restart (None = all).
        No host round trip, so it can follow a step directly (per-environment auto-reset for training loops); the
        draws are those of `reset()` but from the counter-based generator of the device sampler.  The TimeLimit(200)
        bookkeeping of `step()` is one counter for the lock-stepped batch: it restarts only with mask=None; a caller that
        restarts subsets tracks their episode lengths itself (AVG_E_ITERATION in the state record counts them per env)."""
        torch = self.torch
        if seed is None:
            seed = int(self.np_random.randint(1 << 31)) if not hasattr(self, "_device_seed") else self._device_seed
        self._device_seed = seed
        mptr = 0
        if mask is not None:
            self._mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            mptr = self._mask.data_ptr()
        self.sim.reset_device(mptr, seed, self.obs.data_ptr(), self._stream())
        if mask is None:
            self.elapsed = 0
            self._last_reset = "device"
        self.variants = None          # now lives on the device only
        self._needs_reset = False
        return self.obs

    def set_policy(self, blob: bytes):
        """Upload a policy (assistive_vr_gym_b200.policy) for `act()` / `rollout()`."""
        self.sim.upload_policy(blob)
        self._graphs.pop("rollout", None)        # the captured policy launch holds the old blob's arguments
        if not hasattr(self, "actions_dev"):
            self.actions_dev = self.torch.zeros((self.num_envs, self.sim.n_actions), dtype=self.torch.float32, device=self.device)

    def act(self):
        """Deterministic policy action for the current observations (enjoy_vr.py:106-113), on the device."""
        self.sim.policy_act(self.obs.data_ptr(), self.actions_dev.data_ptr(), self._stream())
        return self.actions_dev

    def rollout(self, n_steps: int):
        """`n_steps` of act -> step without leaving the GPU (the loop of enjoy_vr.py:105-117)."""
        out = None
        for _ in range(n_steps):
            if self._needs_reset:
                raise RuntimeError("call reset() before step()")
            self._advance(None)
            out = self._step_result()
        return out

    def set_state(self, env: np.ndarray, variant: Optional[np.ndarray] = None, part: Optional[np.ndarray] = None, settle: int = 0):
        """Import explicit env records ("identical initial states" for parity runs); Feeding / Drinking also
        take the particle records (`part`), and `settle` > 0 runs that many stepSimulation calls before the observation."""
        self.variants = variant
        self.sim.set_state(env, variant)
        if part is not None:
            self.sim.set_particles(part)
        if settle > 0:
            self.sim.settle(settle, 0, self._stream())
        self.sim.reset_obs(self.obs.data_ptr(), self._stream())
        self.elapsed = 0
        self._needs_reset = False

    def get_state(self) -> np.ndarray:
        return self.sim.get_state()

    def get_particles(self) -> np.ndarray:
        """Particle records of Feeding / Drinking (AVG_P_* layout of include/avg_model.h)."""
        return self.sim.get_particles()

    def contact_overflow(self):
        """Per-environment overflow flags accumulated over the episode (AVG_E_OVERFLOW: bit 0 contact points beyond
        AVG_MAX_CONTACT, bit 1 rows, bit 2 broadphase candidates, bit 3 particle contacts, bit 4 particle candidates) as a
        CUDA int32 tensor: a contact the simulator could not keep is flagged, never silently dropped."""
        torch = self.torch
        state = self.sim.get_state_tensor(torch, self.device)
        return state[:, 166].view(torch.int32)

    def _stream(self) -> int:
        return int(self.torch.cuda.current_stream(self.device).cuda_stream)

    def step(self, actions):
        """-> (obs [N, D], reward [N], done [N] bool, info) as CUDA tensors; info follows scratch_itch.py:77."""
        torch = self.torch
        if self._needs_reset:
            raise RuntimeError("call reset() before step()")
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.asarray(actions, dtype=np.float32), device=self.device)
        actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
        if actions.dim() == 1:
            actions = actions.unsqueeze(0)
        if tuple(actions.shape) != (self.num_envs, self.sim.n_actions):
            raise ValueError(f"expected actions of shape {(self.num_envs, self.sim.n_actions)}, got {tuple(actions.shape)}")
        self._advance(actions)
        return self._step_result()

    def _advance(self, actions=None):
        """Device side of step(): eager launches, or one graph replay (actions=None: the policy acts inside the graph)."""
        torch = self.torch
        if self.auto_reset == "device":
            if not hasattr(self, "terminal_obs"):
                self.terminal_obs = torch.empty_like(self.obs)
            if not hasattr(self, "_device_seed"):
                self._device_seed = int(self.np_random.randint(1 << 31))
        if self.cuda_graph and not hasattr(self, "actions_dev"):
            self.actions_dev = torch.zeros((self.num_envs, self.sim.n_actions), dtype=torch.float32, device=self.device)
        if self.cuda_graph and self._eager_steps >= 1:              # first step eager: lazy state (function attributes) settles
            if actions is None:
                self._replay("rollout", True)
            else:
                if actions.data_ptr() != self.actions_dev.data_ptr():
                    self.actions_dev.copy_(actions)
                self._replay("step", False)
            return
        self._eager_steps += 1
        if actions is None:
            actions = self.act()
        self._enqueue_step(actions.data_ptr())

    def _step_result(self):
        torch = self.torch
        if self.auto_reset == "device":
            done = self.done_dev.bool()
            return self.obs, self.reward, done, {
                "total_force_on_human": self.info_dev[:, 0], "task_success": self.info_dev[:, 1].to(torch.int32),
                "action_robot_len": self.action_robot_len, "action_human_len": self.action_human_len,
                "obs_robot_len": self.obs_robot_len, "obs_human_len": self.obs_human_len,
                "TimeLimit.truncated": done, "terminal_observation": self.terminal_obs}
        self.elapsed += 1
        timeout = self.elapsed >= MAX_EPISODE_STEPS                 # gym TimeLimit, __init__.py:21
        done = self.done_dev.bool() | timeout
        info = {"total_force_on_human": self.info_dev[:, 0], "task_success": self.info_dev[:, 1].to(torch.int32),
                "action_robot_len": self.action_robot_len, "action_human_len": self.action_human_len,
                "obs_robot_len": self.obs_robot_len, "obs_human_len": self.obs_human_len,
                "TimeLimit.truncated": timeout, "contact_overflow": self.contact_overflow()}
        if timeout:
            if self.auto_reset:                # gym vector-env convention: the returned observation starts the next episode
                info["terminal_observation"] = self.obs.clone()
                if getattr(self, "_last_reset", "host") == "device":
                    self.reset_device()        # no host round trip (counter-based draws, next episode index)
                else:
                    self.reset()
            else:
                self._needs_reset = False      # like gym, stepping past the limit is the caller's business
        return self.obs, self.reward, done, info

    def step_host(self, actions: np.ndarray):
        """Same step with NumPy buffers through avg_step_host (host<->device copies included)."""
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.num_envs, self.sim.n_actions)
        if not hasattr(self, "_h_obs"):
            # results land in page-locked arrays owned by the env (returned as views, like self.obs on the device)
            self._pinned = [capi.PinnedArray((self.num_envs, self.sim.n_obs), np.float32), capi.PinnedArray((self.num_envs,), np.float32),
                            capi.PinnedArray((self.num_envs,), np.uint8), capi.PinnedArray((self.num_envs, 2), np.float32)]
            self._h_obs, self._h_rew, self._h_done, self._h_info = (p.array for p in self._pinned)
            self._h_success = np.empty(self.num_envs, dtype=np.int32); self._h_doneb = np.empty(self.num_envs, dtype=bool)
        self.sim.step_host(a, self._h_obs, self._h_rew, self._h_done, self._h_info)
        self.elapsed += 1
        timeout = self.elapsed >= MAX_EPISODE_STEPS
        np.copyto(self._h_success, self._h_info[:, 1], casting="unsafe")      # preallocated: no 1.5 MB allocation per step at 393216 envs
        if timeout: self._h_doneb.fill(True)
        else: np.not_equal(self._h_done, 0, out=self._h_doneb)
        info = {"total_force_on_human": self._h_info[:, 0], "task_success": self._h_success,
                "action_robot_len": self.action_robot_len, "action_human_len": self.action_human_len,
                "obs_robot_len": self.obs_robot_len, "obs_human_len": self.obs_human_len}
        return self._h_obs, self._h_rew, self._h_doneb, info

    def pinned_actions(self) -> np.ndarray:
        """A page-locked [N, A] float32 array: actions written here reach the GPU without a staging copy."""
        if not hasattr(self, "_pinned_act"):
            self._pinned_act = capi.PinnedArray((self.num_envs, self.sim.n_actions), np.float32)
        return self._pinned_act.array

    def render(self):
        """env.py:613-620 opens the PyBullet GUI; rendering is out of scope for the batched simulator."""
        return None

    def close(self):
        self.sim.close()


class AssistiveEnvNumpy:
    """`num_envs=1` view with the reference's exact return types: float64 1-D obs, python float reward, bool done,
    info dict of scalars (scratch_itch.py:77-82) — what `examples/random_actions.py` expects."""

    def __init__(self, env_id: str, device: int = 0, seed: int = 1001):
        self.env = BatchedAssistiveEnv(env_id, num_envs=1, device=device, seed=seed)
        self.action_space = self.env.action_space
        self.observation_space = self.env.observation_space

    def seed(self, seed=None):
        return self.env.seed(seed)

    def reset(self):
        return self.env.reset()[0].double().cpu().numpy()

    def step(self, action):
        obs, rew, done, info = self.env.step_host(np.asarray(action, dtype=np.float32)[None])
        out_info = {k: (v if np.isscalar(v) else (float(v[0]) if v.dtype.kind == "f" else int(v[0]))) for k, v in info.items()}
        return obs[0].astype(np.float64), float(rew[0]), bool(done[0]), out_info

    def render(self):
        return None

    def close(self):
        self.env.close()


def make(env_id: str, num_envs: Optional[int] = None, device: int = 0, seed: int = 1001, **kw):
    """`gym.make` replacement.  num_envs=None -> reference-typed single environment; otherwise the batched env."""
    if num_envs is None:
        return AssistiveEnvNumpy(env_id, device=device, seed=seed)
    return BatchedAssistiveEnv(env_id, num_envs=num_envs, device=device, seed=seed, **kw)
