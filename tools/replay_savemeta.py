"""Re-score recorded episodes: the batched simulator's counterpart of the reference's replay_vr_savemeta.py.

    python tools/replay_savemeta.py --replay-dir <dir with participant_*/<run>/{setup.pkl,actions.pkl,frame_0.npz}>

writes observations_vr.pkl = [env_names, observations, rewards, actions, forces, task_success] (same list layout as the
reference, replay_vr_savemeta.py:58-59).  Runs are recorded with assistive_vr_gym_b200.replay.EpisodeRecorder."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200.replay import rescore_directory      # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="replay / re-score recorded episodes")
    ap.add_argument("--replay-dir", required=True)
    ap.add_argument("--out", default="observations_vr.pkl")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args()
    names, obs, rew, act, force, succ = rescore_directory(args.replay_dir, args.out, device=args.device)
    for n, r, f, s in zip(names, rew, force, succ):
        print(n, float(np.sum(r)), float(np.mean(f)), s)         # replay_vr_savemeta.py:55
