"""ORACLE (test infrastructure, not product code): numpy restatement of the reference's vectorised
environment, per-env buffer, mask bookkeeping and rollout driver - same algorithmic structure
(per-env Python loops, Python lists) so that it also serves as the CPU-baseline "port".

Follows (file:line under /root/reference):
  VecMemory / EnvVectorizer / AsyncPPO.worker / run   AsyncTools/AsyncPPO.py:11-33 / :35-102 / :117-146 / :148-165
  the seven utils functions                           AsyncTools/utils.py:3-50
  Memory                                              PPO/Memory.py:7-30
Pinned against the real reference by tests/golden/rollout_*.npz and utils.npz (tests/test_oracle.py).
"""
from __future__ import annotations

import copy

import numpy as np

f32 = np.float32


class FlatMemory:
    """PPO.memory: four flat lists of float32 items."""

    def __init__(self):
        self.states, self.actions, self.rewards, self.dones = [], [], [], []

    def push(self, state, action, reward, done):
        for lst, x in ((self.states, state), (self.actions, action), (self.rewards, reward), (self.dones, done)):
            lst.append(np.asarray(x).astype(f32))

    def clear(self):
        for lst in (self.states, self.actions, self.rewards, self.dones):
            lst.clear()


class PerEnvMemory:
    """VecMemory: one list per env and field; push casts every item to float32."""

    FIELDS = ("states", "actions", "rewards", "dones")

    def __init__(self, num_envs):
        for name in self.FIELDS:
            setattr(self, name, [[] for _ in range(num_envs)])

    def push(self, idx, state, action, reward, done):
        for name, x in zip(self.FIELDS, (state, action, reward, done)):
            getattr(self, name)[idx].append(x.astype(f32))

    def clear(self):
        for name in self.FIELDS:
            for lst in getattr(self, name):
                lst.clear()


# ------------------------------------------------------------------------------- utils.py:3-50
def indexes_of_active(num_envs, terminal):
    return np.flatnonzero(~terminal[:num_envs]) if len(terminal) == num_envs else np.arange(num_envs)[~terminal]


def number_of_active(terminal):
    return np.count_nonzero(~terminal)


def range_of_active(terminal):
    return np.arange(number_of_active(terminal))


def states_dropout(states, dones):
    return states[~dones]


def buffer_append(buffer, states, actions, rewards, dones, terminal, num_envs):
    for rank, env in enumerate(indexes_of_active(num_envs, terminal)):
        buffer.push(env, states[rank], actions[rank], rewards[rank], dones[rank])


def update_mask(terminal, dones):
    terminal[np.flatnonzero(~terminal)] = dones
    return terminal


def transfer(buffer, target):
    for name in PerEnvMemory.FIELDS:
        dst = getattr(target, name)
        for per_env in getattr(buffer, name):  # env-major, time-minor
            dst.extend(per_env)
    buffer.clear()


# ------------------------------------------------------------------------------- EnvVectorizer
class Vectorizer:
    def __init__(self, env, num_envs=1):
        self.envs = [copy.deepcopy(env) for _ in range(num_envs)]
        self.envs_active = np.zeros(num_envs, bool)  # True = terminal
        self.num_envs = num_envs
        self.action_space, self.observation_space = env.action_space, env.observation_space

    def reset(self):
        out = [e.reset() for e in self.envs]
        self.envs_active = np.zeros(self.num_envs, bool)
        return np.stack([o for o, _ in out], 0), [i for _, i in out]

    def step(self, actions):
        res = [self.envs[e].step(actions[rank]) for rank, e in enumerate(np.flatnonzero(~self.envs_active))]
        cols = list(zip(*res))
        return tuple(np.stack(c, 0) for c in cols)

    def close(self):
        pass


def worker(vec, buffer, policy_act, memory):
    """One episode per env (AsyncPPO.worker). `policy_act(states f32 [n,O]) -> actions [n]/[n,A]`.
    Returns (reward_score, step_score)."""
    states = vec.reset()[0]
    reward_score, step_score = 0.0, 0
    while True:
        actions = policy_act(states)
        nxt, rewards, dones, truncs, _ = vec.step(actions)
        fin = dones | truncs
        buffer_append(buffer, states, actions, rewards, fin, vec.envs_active, vec.num_envs)
        reward_score += np.sum(rewards)
        step_score += int(np.sum(~vec.envs_active))
        states = states_dropout(nxt, fin)
        vec.envs_active = update_mask(vec.envs_active, fin)
        if np.all(vec.envs_active):
            transfer(buffer, memory)
            return reward_score, step_score
