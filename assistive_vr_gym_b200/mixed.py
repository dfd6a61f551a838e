"""Mixed-task batches (BASELINE.json configs[3]: "mixed-task batch (all robots)").

Environments of one id share a model, so a batch is homogeneous by construction; a mixed batch is one homogeneous
sub-batch per (task, robot[, human-active]) id.  Each sub-batch is its own handle (own arena, own slot of the constant-memory
model table); `MixedBatch` steps them on separate CUDA streams so that their kernels overlap on the device, and joins them
back into the caller's stream.  Across GPUs whole sub-batches (or env ranges of each, `sharding.py`) go to different
ranks: there is no exchange between sub-batches either way.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

from .envs import BatchedAssistiveEnv, REGISTRY


class MixedBatch:
    def __init__(self, env_ids: Optional[Sequence[str]] = None, envs_per_id: int = 4096, device: int = 0, seed: int = 1001, **kw):
        import torch
        self.torch = torch
        self.env_ids: List[str] = list(env_ids) if env_ids is not None else sorted(REGISTRY)
        self.envs: Dict[str, BatchedAssistiveEnv] = {i: BatchedAssistiveEnv(i, num_envs=envs_per_id, device=device, seed=seed + k, **kw)
                                                     for k, i in enumerate(self.env_ids)}
        self.device = torch.device("cuda", device)
        self.streams = {i: torch.cuda.Stream(device=self.device) for i in self.env_ids}
        self.num_envs = envs_per_id * len(self.env_ids)

    def _fan_out(self, fn):
        torch = self.torch
        cur = torch.cuda.current_stream(self.device)
        out = {}
        for i in self.env_ids:
            s = self.streams[i]
            s.wait_stream(cur)                         # inputs produced on the caller's stream are visible
            with torch.cuda.stream(s):
                out[i] = fn(i, self.envs[i])
        for i in self.env_ids:
            cur.wait_stream(self.streams[i])           # results are visible to the caller's stream
        return out

    def reset(self):
        return {i: e.reset() for i, e in self.envs.items()}

    def reset_device(self, seed: Optional[int] = None):
        return self._fan_out(lambda i, e: e.reset_device(seed=seed))

    def step(self, actions: Dict[str, object]):
        """actions: {env_id: tensor [envs_per_id, n_actions(env_id)]} -> {env_id: (obs, reward, done, info)}"""
        return self._fan_out(lambda i, e: e.step(actions[i]))

    def sample_actions(self, generator=None):
        torch = self.torch
        return {i: torch.rand((e.num_envs, e.sim.n_actions), device=self.device, generator=generator) * 2 - 1 for i, e in self.envs.items()}

    def close(self):
        for e in self.envs.values():
            e.close()
