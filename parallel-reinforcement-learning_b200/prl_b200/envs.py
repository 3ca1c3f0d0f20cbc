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
    # state = {position, velocity, stepped}: gymnasium's state array is float64 until the first step and float32 afterwards
    "MountainCarContinuous-v0": dict(S=3, O=2, A=1, continuous=True, max_steps=999, low=-1.0, high=1.0),
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
            self.action_space = SimpleNamespace(shape=(s["A"],), dtype=np.float32, low=s.get("low", -2.0), high=s.get("high", 2.0))
        else:
            self.action_space = SimpleNamespace(n=s["A"], shape=(), dtype=np.int64)
        self.is_continuous = s["continuous"]
        self.observ_dim, self.action_dim, self.state_dim = s["O"], s["A"], s["S"]

    def close(self):
        pass

    # ---- the gymnasium single-env surface, for playback loops like the reference's Test.py:19-33 (`state, _ = env.reset()`,
    # `env.step(action)`): ONE env on the device, stepped by the same kernels as the vectorised path (batch 1) -----------------
    def _single(self):
        if getattr(self, "_sim1", None) is None:
            import torch

            from . import ops

            self._sim1 = ops.EnvState(self.env_id, 1, self.max_episode_steps)
            self._seed1 = int(torch.randint(0, 2 ** 62, (1,)).item())
            self._episode1, self._numpy1 = 0, False
            self._idx1 = torch.zeros(1, dtype=torch.int32, device=self._sim1.state.device)
        return self._sim1

    def reset(self, seed=None, options=None):
        """-> (observation float32 [O], info).  seed: gymnasium's env.reset(seed=...) stream (numpy PCG64, bit for bit)."""
        import torch

        sim = self._single()
        self._episode1 += 1
        if seed is not None:
            self._numpy1 = True
            obs = sim.reset_numpy(torch.from_numpy(np.asarray([seed], dtype=np.uint64).view(np.int64)).to(sim.state.device))
        elif self._numpy1:
            obs = sim.reset_numpy()
        else:
            obs = sim.reset(self._seed1, self._episode1)
        return obs[0].cpu().numpy(), {}

    def step(self, action):
        """-> (observation, reward, terminated, truncated, info) like gymnasium (TimeLimit included)."""
        import torch

        sim = self._single()
        a = np.asarray(action)
        if self.is_continuous:
            a_dev = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32).reshape(1, -1)).to(sim.state.device)
        else:
            a_dev = torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64).reshape(1)).to(sim.state.device)
        obs, reward, done, trunc = sim.step(self._idx1, 1, a_dev)
        return obs[0].cpu().numpy(), float(reward[0].item()), bool(done[0].item()), bool(trunc[0].item()), {}

    def render(self):
        return None

    def __repr__(self):
        return f"EnvDescriptor({self.env_id!r}, max_episode_steps={self.max_episode_steps})"


def make(env_id: str, max_episode_steps: int | None = None, **_ignored) -> EnvDescriptor:
    return EnvDescriptor(env_id, max_episode_steps)


def play(ppo, env, episodes: int = 1, seed=None, on_step=None):
    """The reference's Test.py loop (Test.py:19-33) as a helper: batch-1 `ppo.get_action` + `env.step` until done | truncate,
    `episodes` times.  Returns the list of episode rewards.  on_step(state, action, reward, done) is called after every step
    (Test.py renders and updates a progress bar there)."""
    import torch

    totals = []
    for ep in range(int(episodes)):
        state, _ = env.reset(seed=None if seed is None else seed + ep)
        total = 0.0
        while True:
            action = ppo.get_action(torch.from_numpy(state).unsqueeze(0))
            state, reward, done, truncate, _ = env.step(action.squeeze(0))
            total += reward
            if on_step is not None:
                on_step(state, action, reward, done or truncate)
            if done or truncate:
                break
        totals.append(total)
    return totals


def describe(env) -> EnvDescriptor:
    """Accept our descriptor or anything gym-like exposing `.spec.id` (a real gymnasium env)."""
    if isinstance(env, EnvDescriptor):
        return env
    spec = getattr(env, "spec", None)
    env_id = getattr(spec, "id", None)
    if env_id is None:
        raise TypeError("EnvVectorizer needs an env with .spec.id (gym.make(...)) or prl_b200.make(...)")
    return EnvDescriptor(env_id, getattr(spec, "max_episode_steps", None))
