"""B200-native batched simulator behind the assistive_gym API (ScratchItch path)."""
from .envs import make, BatchedAssistiveEnv, AssistiveEnvNumpy, REGISTRY  # noqa: F401
