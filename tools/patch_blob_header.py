"""Set header fields of the compiled ModelBlobs under assistive_vr_gym_b200/data/ in place (fields that do not change the layout,
e.g. the warm-starting factor added in round 2, without re-running the model compiler and its IK / TOC searches).
usage: python tools/patch_blob_header.py warmstart=0.1 [file.npz ...]"""
import glob, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from assistive_vr_gym_b200.compiler.blob import HEADER_DT
sets = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
files = [a for a in sys.argv[1:] if "=" not in a] or sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assistive_vr_gym_b200", "data", "*.npz")))
for f in files:
    z = np.load(f)
    out = {}
    n = 0
    for k in z.files:
        a = z[k]
        if k.startswith("blob_"):
            b = bytearray(a.tobytes())
            h = np.frombuffer(b, dtype=HEADER_DT, count=1)
            for name, v in sets.items():
                h[name] = float(v) if HEADER_DT[name].kind == "f" else int(v)
            a = np.frombuffer(bytes(b), dtype=np.uint8); n += 1
        out[k] = a
    np.savez_compressed(f, **out)
    print(os.path.basename(f), n, "blobs patched")
