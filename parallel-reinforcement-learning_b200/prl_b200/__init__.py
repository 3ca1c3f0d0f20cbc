"""prl_b200 - host side of the B200-native data-parallel PPO hot path.

  prl_b200._lib   ctypes binding of libprl_b200.so (C ABI, include/prl_b200.h); no CPU fallback
  prl_b200.ops    one torch-tensor wrapper per C entry point
  prl_b200.envs   env descriptors (`make("CartPole-v1")`) standing in for gym.make
The drop-in packages `PPO` and `AsyncTools` (same import paths as the reference) sit next to this package.
"""
from ._lib import PrlError, env_info, load_library  # noqa: F401
from .envs import make, play  # noqa: F401

__all__ = ["PrlError", "env_info", "load_library", "make", "play"]
