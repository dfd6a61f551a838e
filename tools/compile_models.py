"""Compile the reference assets into ModelBlobs + reset pools and store them under assistive_vr_gym_b200/data/.

Run in the build container (needs /root/reference); the outputs are derived data (hull vertices, inertias, IK start
poses) that travel to the GPU box, where the reference tree does not exist.

    python tools/compile_models.py [--assets /root/reference/assistive_gym/envs/assets] [--pool 64]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200.compiler.blob import scene_to_blob          # noqa: E402
from assistive_vr_gym_b200.compiler.reset import build_reset_data      # noqa: E402
from assistive_vr_gym_b200.compiler.scene import build_scratch_itch    # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--assets", default="/root/reference/assistive_gym/envs/assets")
    ap.add_argument("--pool", type=int, default=64)
    args = ap.parse_args()
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assistive_vr_gym_b200", "data")
    os.makedirs(out_dir, exist_ok=True)
    for human_control in (False, True):
        payload = {}
        rng = np.random.RandomState(1001)              # env.py:53 default seed
        for v, gender in enumerate(("male", "female")):
            scene = build_scratch_itch(args.assets, "jaco", gender, human_control=human_control)
            blob = scene_to_blob(scene)
            payload[f"blob_{v}"] = np.frombuffer(blob, dtype=np.uint8)
            rd = build_reset_data(scene, rng, ik_pool=args.pool)
            for k, a in rd.items():
                payload[f"reset_{v}_{k}"] = a
            print(gender, "human_control" if human_control else "", scene.info["n_pairs"], "pairs", len(blob), "bytes",
                  {k: round(e, 5) for k, e in scene.info["hull_errors"].items()})
        name = "ScratchItchJaco" + ("Human" if human_control else "") + ".npz"
        np.savez_compressed(os.path.join(out_dir, name), **payload)
        print("wrote", name)
