"""Env descriptors: what `gym.make(id)` hands to the reference (train.py:8-13) - an object carrying the env id,
observation/action space shapes and the TimeLimit - without any physics on the host.  The physics lives in the
CUDA kernels (csrc/envs.cuh); EnvVectorizer maps the descriptor (or a real gymnasium env's `spec.id`) to them.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np

_SPECS = {
    "CartPole-v1": dict(S=4, O=4, A=2, continuous=False, max_steps=500),
    "Pendulum-v1": dict(S=2, O=3, A=1, continuous=True, max_steps=200),
    "Acrobot-v1": dict(S=4, O=6, A=3, continuous=False, max_steps=500),
    "MountainCar-v0": dict(S=2, O=2, A=3, continuous=False, max_steps=200),
}


class EnvDescriptor:
    """Stands in for a gymnasium env object at the API boundary (AsyncPPO(env=...), EnvVectorizer(env=...))."""

    def __init__(self, env_id: str, max_episode_steps: int | None = None):
        if env_id not in _SPECS:
            raise ValueError(f"unsupported env id {env_id!r}; kernels exist for {sorted(_SPECS)}")
        s = _SPECS[env_id]
        self.env_id = env_id
        self.max_episode_steps = int(max_episode_steps or s["max_steps"])
        self.spec = SimpleNamespace(id=env_id, max_episode_steps=self.max_episode_steps)
        self.observation_space = SimpleNamespace(shape=(s["O"],), dtype=np.float32)
        if s["continuous"]:
            self.action_space = SimpleNamespace(shape=(s["A"],), dtype=np.float32, low=-2.0, high=2.0)
        else:
            self.action_space = SimpleNamespace(n=s["A"], shape=(), dtype=np.int64)
        self.is_continuous = s["continuous"]
        self.observ_dim, self.action_dim, self.state_dim = s["O"], s["A"], s["S"]

    def close(self):
        pass

    def __repr__(self):
        return f"EnvDescriptor({self.env_id!r}, max_episode_steps={self.max_episode_steps})"


def make(env_id: str, max_episode_steps: int | None = None, **_ignored) -> EnvDescriptor:
    return EnvDescriptor(env_id, max_episode_steps)


def describe(env) -> EnvDescriptor:
    """Accept our descriptor or anything gym-like exposing `.spec.id` (a real gymnasium env)."""
    if isinstance(env, EnvDescriptor):
        return env
    spec = getattr(env, "spec", None)
    env_id = getattr(spec, "id", None)
    if env_id is None:
        raise TypeError("EnvVectorizer needs an env with .spec.id (gym.make(...)) or prl_b200.make(...)")
    return EnvDescriptor(env_id, getattr(spec, "max_episode_steps", None))
