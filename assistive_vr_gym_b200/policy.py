"""Policies for on-device rollouts (reference enjoy_vr.py:77-117).

The reference evaluates pickled `a2c_ppo_acktr` checkpoints `(actor_critic, ob_rms)`; that package is not part of the
reference tree and no checkpoint ships with it (`trained_models/ppo/` is empty), so this module packs the arrays of the
default MLP actor -- from a state dict when one is available, or synthetic ones for throughput runs -- into the blob
`avg_upload_policy` takes (include/avg_model.h, AvgPolicyHeader)."""
from __future__ import annotations

import struct

import numpy as np

POLICY_MAGIC = 0x4C505641
HIDDEN = 64


def pack_policy(ob_mean, ob_var, W1, b1, W2, b2, W3, b3, clip_obs: float = 10.0, eps: float = 1e-8) -> bytes:
    """W1 [n_in, 64], W2 [64, 64], W3 [64, n_out] (input-major, i.e. torch `Linear.weight.T`)."""
    W1 = np.asarray(W1, dtype=np.float32); W2 = np.asarray(W2, dtype=np.float32); W3 = np.asarray(W3, dtype=np.float32)
    n_in, n_out = W1.shape[0], W3.shape[1]
    assert W1.shape == (n_in, HIDDEN) and W2.shape == (HIDDEN, HIDDEN) and W3.shape == (HIDDEN, n_out)
    parts = [np.asarray(ob_mean, dtype=np.float32).reshape(n_in), np.asarray(ob_var, dtype=np.float32).reshape(n_in), W1.ravel(),
             np.asarray(b1, dtype=np.float32).reshape(HIDDEN), W2.ravel(), np.asarray(b2, dtype=np.float32).reshape(HIDDEN), W3.ravel(),
             np.asarray(b3, dtype=np.float32).reshape(n_out)]
    return struct.pack("<IiiffIII", POLICY_MAGIC, n_in, n_out, clip_obs, eps, 0, 0, 0) + b"".join(p.tobytes() for p in parts)


def from_state_dict(sd: dict, ob_mean, ob_var, **kw) -> bytes:
    """a2c_ppo_acktr `Policy.state_dict()` (MLPBase actor + DiagGaussian mean) and VecNormalize's ob_rms."""
    g = lambda k: np.asarray(sd[k].detach().cpu().numpy() if hasattr(sd[k], "detach") else sd[k], dtype=np.float32)
    return pack_policy(ob_mean, ob_var, g("base.actor.0.weight").T, g("base.actor.0.bias"), g("base.actor.2.weight").T, g("base.actor.2.bias"),
                       g("dist.fc_mean.weight").T, g("dist.fc_mean.bias"), **kw)


def synthetic_policy(n_in: int, n_out: int, seed: int = 0, gain: float = 1.0):
    """Orthogonal-initialised actor (a2c_ppo_acktr initialises with orthogonal weights, gain sqrt(2) / 0.01 on the mean
    layer; here the mean layer keeps `gain` so the arm actually moves) and unit observation statistics.
    -> (blob, arrays dict) -- "synthetic-policy" in every report that uses it."""
    rng = np.random.RandomState(seed)

    def ortho(rows, cols, g):
        a = rng.normal(size=(max(rows, cols), min(rows, cols)))
        q, _ = np.linalg.qr(a)
        q = q if rows >= cols else q.T
        return (g * q[:rows, :cols]).astype(np.float32)

    arrs = dict(ob_mean=np.zeros(n_in, np.float32), ob_var=np.ones(n_in, np.float32),
                W1=ortho(n_in, HIDDEN, np.sqrt(2)), b1=np.zeros(HIDDEN, np.float32), W2=ortho(HIDDEN, HIDDEN, np.sqrt(2)),
                b2=np.zeros(HIDDEN, np.float32), W3=ortho(HIDDEN, n_out, gain), b3=np.zeros(n_out, np.float32))
    return pack_policy(**arrs), arrs
