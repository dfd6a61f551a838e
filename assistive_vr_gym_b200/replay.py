"""Episode logs and replay / re-scoring.

The reference keeps three artefacts per recorded episode (VR study runs, `scratch_itch.py:47-51,140-144,266-272`):
`setup.pkl` = `[robot_type, gender, hipbone_to_mouth_height]`, `actions.pkl` = the list of actions of the 200 steps, and
`frame_%d.bullet` PyBullet snapshots (`p.saveBullet`, frame 0 = the post-reset world).  `replay_vr_savemeta.py:17-59`
walks `participant_*/<run>` directories, replays every run (`env.replay_setup(dir)`, `env.reset()`, `env.step(...)` until
done) and pickles `[env_names, observations, rewards, actions, forces, task_success]` into `observations_vr.pkl`.

Here the two pickles keep the reference's exact layout; the snapshot is our own state record (`frame_0.npz`: env record +
model variant) because `.bullet` files are PyBullet-internal.  A replay re-simulates the logged actions from the snapshot
(the kernels are deterministic, so it reproduces the recorded rewards bit for bit), which is what the reference's replay
does with `frame_%d.bullet` + `action_list` (`scratch_itch.py:33-41`).
"""
from __future__ import annotations

import glob
import os
import pickle
from typing import List, Optional

import numpy as np

_H2M = {"male": 0.6, "female": 0.54}          # scratch_itch.py:161, bed_bathing.py:196


def _gender_of(env, variant: int) -> str:
    per_gender = max(len(env.blobs) // 2, 1)
    return "male" if variant // per_gender == 0 else "female"


class EpisodeRecorder:
    """Logs one environment of a batch.  `start()` right after `reset()`, `record(actions)` before every `step(actions)`,
    `close()` at the end of the episode (the reference writes `actions.pkl` at iteration 200, scratch_itch.py:47-51)."""

    def __init__(self, env, directory: str, env_index: int = 0):
        self.env, self.directory, self.k = env, directory, int(env_index)
        self.action_list: List[np.ndarray] = []
        os.makedirs(directory, exist_ok=True)

    def start(self) -> None:
        st = self.env.get_state()
        variant = int(self.env.variants[self.k]) if self.env.variants is not None else int(self.env.sim.get_variants()[self.k])
        gender = _gender_of(self.env, variant)
        with open(os.path.join(self.directory, "setup.pkl"), "wb") as f:                       # scratch_itch.py:269-272
            pickle.dump([self.env.spec["robot"], gender, _H2M[gender]], f)
        np.savez(os.path.join(self.directory, "frame_0.npz"), env_id=self.env.env_id, record=st[self.k], variant=variant)
        self.action_list = []

    def record(self, actions) -> None:
        a = actions.detach().cpu().numpy() if hasattr(actions, "detach") else np.asarray(actions)
        self.action_list.append(np.asarray(a.reshape(self.env.num_envs, -1)[self.k], dtype=np.float32).copy())

    def close(self) -> None:
        with open(os.path.join(self.directory, "actions.pkl"), "wb") as f:                     # scratch_itch.py:50-51
            pickle.dump(self.action_list, f)


def load_episode(directory: str):
    """-> (env_id, record [192] float32, variant, [robot_type, gender, hipbone_to_mouth_height], action_list)"""
    with open(os.path.join(directory, "setup.pkl"), "rb") as f:
        setup = pickle.load(f)                                                                 # scratch_itch.py:140-144
    with open(os.path.join(directory, "actions.pkl"), "rb") as f:
        action_list = pickle.load(f)
    z = np.load(os.path.join(directory, "frame_0.npz"))
    return str(z["env_id"]), np.asarray(z["record"], dtype=np.float32), int(z["variant"]), setup, action_list


def replay(directory: str, device: int = 0, make=None) -> dict:
    """The loop of replay_vr_savemeta.py:25-47 for one recorded run: observations, rewards, forces and the final
    task_success of the logged actions re-simulated from the snapshot."""
    if make is None:
        from .envs import make
    env_id, record, variant, setup, action_list = load_episode(directory)
    env = make(env_id, num_envs=1, device=device)
    env.set_state(record[None, :], np.asarray([variant], dtype=np.int32))
    observations, rewards, forces = [], [], []
    task_success = 0
    for a in action_list:
        obs, rew, done, info = env.step_host(np.asarray(a, dtype=np.float32)[None, :])
        observations.append(obs[0].astype(np.float64)); rewards.append(float(rew[0]))
        forces.append(float(info["total_force_on_human"][0])); task_success = int(info["task_success"][0])
    env.close()
    return dict(env_id=env_id, setup=setup, observations=observations, rewards=rewards, actions=action_list, forces=forces,
                task_success=task_success)


def rescore_directory(replay_dir: str, out: Optional[str] = "observations_vr.pkl", device: int = 0, make=None):
    """replay_vr_savemeta.py: every `participant_*/<run>` under `replay_dir` -> one pickle with the reference's list layout
    `[env_names, observations, rewards, actions, forces, task_success]`."""
    names, obs_all, rew_all, act_all, force_all, succ_all = [], [], [], [], [], []
    for d in sorted(glob.glob(os.path.join(replay_dir, "participant_*", "*"))):
        if not os.path.exists(os.path.join(d, "actions.pkl")):
            continue
        r = replay(d, device=device, make=make)
        names.append(d); obs_all.append(r["observations"]); rew_all.append(r["rewards"]); act_all.append(r["actions"])
        force_all.append(r["forces"]); succ_all.append(r["task_success"])
    result = [names, obs_all, rew_all, act_all, force_all, succ_all]
    if out:
        with open(out, "wb") as f:
            pickle.dump(result, f, pickle.HIGHEST_PROTOCOL)
    return result
