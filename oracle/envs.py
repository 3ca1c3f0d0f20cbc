"""ORACLE (test infrastructure, not product code): numpy restatement of the env physics.

The reference steps `gymnasium==1.1.1` envs one at a time (`/root/reference/AsyncTools/AsyncPPO.py:73-78`,
`envs[env_idx].step(actions[i])`; reset at `:53`).  gymnasium is a third-party dependency that is NOT
vendored under `/root/reference`, is not installed in this image and cannot be installed (no network).
This file restates its *published* classic-control algorithms from memory:

  gymnasium/envs/classic_control/cartpole.py   (CartPoleEnv.step/reset)
  gymnasium/envs/classic_control/pendulum.py   (PendulumEnv.step/reset, angle_normalize)
  gymnasium/envs/classic_control/acrobot.py    (AcrobotEnv.step/_dsdt, rk4, wrap, bound)
  gymnasium/envs/classic_control/mountain_car.py (MountainCarEnv.step/reset)
  gymnasium/envs/classic_control/continuous_mountain_car.py (Continuous_MountainCarEnv.step/reset)
  gymnasium/wrappers/common.py                 (TimeLimit)

PARITY UNPINNED against gymnasium itself: there is no gymnasium source, wheel or golden vector to
check this restatement against.  What *is* pinned is everything around it: the reference's own
EnvVectorizer / VecMemory / utils / AsyncPPO.worker code is run verbatim over these env objects
(tests/golden/gen_golden.py) and the product is compared with that.

The arithmetic is written with numpy scalars exactly the way gymnasium writes it, so that the
numpy-2 (NEP 50) promotion rules, `np.sin/np.cos` (= glibc libm here) and `np.float64.__pow__`
(= libm `pow`, which is NOT always `x*x`) are what a gymnasium run on this box would execute.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may
import this module.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

__all__ = ["make", "CartPole", "Pendulum", "Acrobot", "MountainCar", "MountainCarContinuous", "ENV_IDS"]


class _Space(SimpleNamespace):
    pass


class _BaseEnv:
    """Common plumbing: TimeLimit (gymnasium/wrappers/common.py::TimeLimit), lazy per-env RNG, state
    injection for teacher-forced runs."""

    env_id = ""
    default_max_episode_steps = 0

    def __init__(self, max_episode_steps: int | None = None):
        self.max_episode_steps = int(max_episode_steps or self.default_max_episode_steps)
        self._elapsed_steps = 0
        self._np_random = None  # lazily seeded from OS entropy, as gymnasium does
        self._injected = None
        self.state = None
        self.spec = SimpleNamespace(id=self.env_id, max_episode_steps=self.max_episode_steps)

    # -- RNG ---------------------------------------------------------------------------------------
    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.default_rng()
        return self._np_random

    def inject_state(self, state):
        """Teacher forcing: the next reset() installs `state` instead of drawing from the RNG."""
        self._injected = np.array(state, dtype=np.float64)

    # -- gym API -----------------------------------------------------------------------------------
    def reset(self, seed=None, options=None):
        if seed is not None:
            self._np_random = np.random.default_rng(seed)
        self._elapsed_steps = 0
        if self._injected is not None:
            self.state = self._injected
            self._injected = None
        else:
            self.state = self._draw_state()
        return self._obs(), {}

    def step(self, action):
        obs, reward, terminated = self._physics(action)
        # TimeLimit: elapsed += 1; truncated = elapsed >= max_episode_steps
        self._elapsed_steps += 1
        truncated = self._elapsed_steps >= self.max_episode_steps
        return obs, reward, terminated, truncated, {}

    def close(self):
        pass


class CartPole(_BaseEnv):
    env_id = "CartPole-v1"
    default_max_episode_steps = 500

    def __init__(self, max_episode_steps=None):
        super().__init__(max_episode_steps)
        self.gravity = 9.8
        self.masscart = 1.0
        self.masspole = 0.1
        self.total_mass = self.masspole + self.masscart
        self.length = 0.5
        self.polemass_length = self.masspole * self.length
        self.force_mag = 10.0
        self.tau = 0.02
        self.theta_threshold_radians = 12 * 2 * math.pi / 360
        self.x_threshold = 2.4
        self.observation_space = _Space(shape=(4,), dtype=np.float32)
        self.action_space = _Space(n=2, shape=(), dtype=np.int64)

    def _draw_state(self):
        return self.np_random.uniform(low=-0.05, high=0.05, size=(4,))

    def _obs(self):
        return np.array(self.state, dtype=np.float32)

    def _physics(self, action):
        x, x_dot, theta, theta_dot = self.state
        force = self.force_mag if action == 1 else -self.force_mag
        costheta = np.cos(theta)
        sintheta = np.sin(theta)
        temp = (force + self.polemass_length * np.square(theta_dot) * sintheta) / self.total_mass
        thetaacc = (self.gravity * sintheta - costheta * temp) / (
            self.length * (4.0 / 3.0 - self.masspole * np.square(costheta) / self.total_mass)
        )
        xacc = temp - self.polemass_length * thetaacc * costheta / self.total_mass
        # explicit Euler, every update from the OLD values
        x = x + self.tau * x_dot
        x_dot = x_dot + self.tau * xacc
        theta = theta + self.tau * theta_dot
        theta_dot = theta_dot + self.tau * thetaacc
        self.state = np.array((x, x_dot, theta, theta_dot), dtype=np.float64)
        terminated = bool(
            x < -self.x_threshold
            or x > self.x_threshold
            or theta < -self.theta_threshold_radians
            or theta > self.theta_threshold_radians
        )
        return self._obs(), 1.0, terminated


def _angle_normalize(x):
    return ((x + np.pi) % (2 * np.pi)) - np.pi


class Pendulum(_BaseEnv):
    env_id = "Pendulum-v1"
    default_max_episode_steps = 200

    def __init__(self, max_episode_steps=None, g=10.0):
        super().__init__(max_episode_steps)
        self.max_speed = 8
        self.max_torque = 2.0
        self.dt = 0.05
        self.g = g
        self.m = 1.0
        self.l = 1.0
        self.observation_space = _Space(shape=(3,), dtype=np.float32)
        self.action_space = _Space(shape=(1,), dtype=np.float32, low=-2.0, high=2.0)

    def _draw_state(self):
        high = np.array([np.pi, 1.0])
        return self.np_random.uniform(low=-high, high=high)

    def _obs(self):
        theta, thetadot = self.state
        return np.array([np.cos(theta), np.sin(theta), thetadot], dtype=np.float32)

    def _physics(self, u):
        th, thdot = self.state
        g, m, l, dt = self.g, self.m, self.l, self.dt
        # the action arrives as a float32 array of shape (1,); clip keeps float32 (NEP 50)
        u = np.clip(u, -self.max_torque, self.max_torque)[0]
        costs = _angle_normalize(th) ** 2 + 0.1 * thdot**2 + 0.001 * (u**2)
        newthdot = thdot + (3 * g / (2 * l) * np.sin(th) + 3.0 / (m * l**2) * u) * dt
        newthdot = np.clip(newthdot, -self.max_speed, self.max_speed)
        newth = th + newthdot * dt
        self.state = np.array([newth, newthdot])
        return self._obs(), -costs, False


def _wrap(x, m, M):
    diff = M - m
    while x > M:
        x = x - diff
    while x < m:
        x = x + diff
    return x


def _bound(x, m, M):
    return min(max(x, m), M)


def _rk4(derivs, y0, t):
    yout = np.zeros((len(t), len(y0)), np.float64)
    yout[0] = y0
    for i in np.arange(len(t) - 1):
        this = t[i]
        dt = t[i + 1] - this
        dt2 = dt / 2.0
        y0 = yout[i]
        k1 = np.asarray(derivs(y0))
        k2 = np.asarray(derivs(y0 + dt2 * k1))
        k3 = np.asarray(derivs(y0 + dt2 * k2))
        k4 = np.asarray(derivs(y0 + dt * k3))
        yout[i + 1] = y0 + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    return yout[-1][:4]


class Acrobot(_BaseEnv):
    env_id = "Acrobot-v1"
    default_max_episode_steps = 500

    dt = 0.2
    LINK_LENGTH_1 = 1.0
    LINK_LENGTH_2 = 1.0
    LINK_MASS_1 = 1.0
    LINK_MASS_2 = 1.0
    LINK_COM_POS_1 = 0.5
    LINK_COM_POS_2 = 0.5
    LINK_MOI = 1.0
    MAX_VEL_1 = 4 * np.pi
    MAX_VEL_2 = 9 * np.pi
    AVAIL_TORQUE = [-1.0, 0.0, +1]

    def __init__(self, max_episode_steps=None):
        super().__init__(max_episode_steps)
        self.observation_space = _Space(shape=(6,), dtype=np.float32)
        self.action_space = _Space(n=3, shape=(), dtype=np.int64)

    def _draw_state(self):
        # gymnasium casts the freshly drawn state to float32
        return self.np_random.uniform(low=-0.1, high=0.1, size=(4,)).astype(np.float32)

    def _obs(self):
        s = self.state
        return np.array(
            [np.cos(s[0]), np.sin(s[0]), np.cos(s[1]), np.sin(s[1]), s[2], s[3]], dtype=np.float32
        )

    def _dsdt(self, s_augmented):
        cos, sin, pi = np.cos, np.sin, np.pi
        m1, m2 = self.LINK_MASS_1, self.LINK_MASS_2
        l1 = self.LINK_LENGTH_1
        lc1, lc2 = self.LINK_COM_POS_1, self.LINK_COM_POS_2
        I1 = I2 = self.LINK_MOI
        g = 9.8
        a = s_augmented[-1]
        s = s_augmented[:-1]
        theta1, theta2, dtheta1, dtheta2 = s[0], s[1], s[2], s[3]
        d1 = m1 * lc1**2 + m2 * (l1**2 + lc2**2 + 2 * l1 * lc2 * cos(theta2)) + I1 + I2
        d2 = m2 * (lc2**2 + l1 * lc2 * cos(theta2)) + I2
        phi2 = m2 * lc2 * g * cos(theta1 + theta2 - pi / 2.0)
        phi1 = (
            -m2 * l1 * lc2 * dtheta2**2 * sin(theta2)
            - 2 * m2 * l1 * lc2 * dtheta2 * dtheta1 * sin(theta2)
            + (m1 * lc1 + m2 * l1) * g * cos(theta1 - pi / 2)
            + phi2
        )
        # "book" variant
        ddtheta2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * dtheta1**2 * sin(theta2) - phi2) / (
            m2 * lc2**2 + I2 - d2**2 / d1
        )
        ddtheta1 = -(d2 * ddtheta2 + phi1) / d1
        return dtheta1, dtheta2, ddtheta1, ddtheta2, 0.0

    def _physics(self, a):
        s = self.state
        torque = self.AVAIL_TORQUE[a]
        s_augmented = np.append(s, torque)
        ns = _rk4(self._dsdt, s_augmented, [0, self.dt])
        ns[0] = _wrap(ns[0], -np.pi, np.pi)
        ns[1] = _wrap(ns[1], -np.pi, np.pi)
        ns[2] = _bound(ns[2], -self.MAX_VEL_1, self.MAX_VEL_1)
        ns[3] = _bound(ns[3], -self.MAX_VEL_2, self.MAX_VEL_2)
        self.state = ns
        terminated = bool(-np.cos(ns[0]) - np.cos(ns[1] + ns[0]) > 1.0)
        reward = -1.0 if not terminated else 0.0
        return self._obs(), reward, terminated


class MountainCar(_BaseEnv):
    """gymnasium/envs/classic_control/mountain_car.py::MountainCarEnv (restated from memory, like the others)."""

    env_id = "MountainCar-v0"
    default_max_episode_steps = 200

    def __init__(self, max_episode_steps=None, goal_velocity=0):
        super().__init__(max_episode_steps)
        self.min_position = -1.2
        self.max_position = 0.6
        self.max_speed = 0.07
        self.goal_position = 0.5
        self.goal_velocity = goal_velocity
        self.force = 0.001
        self.gravity = 0.0025
        self.observation_space = _Space(shape=(2,), dtype=np.float32)
        self.action_space = _Space(n=3, shape=(), dtype=np.int64)

    def _draw_state(self):
        return np.array([self.np_random.uniform(low=-0.6, high=-0.4), 0])

    def _obs(self):
        return np.array(self.state, dtype=np.float32)

    def _physics(self, action):
        position, velocity = self.state
        velocity += (action - 1) * self.force + math.cos(3 * position) * (-self.gravity)
        velocity = np.clip(velocity, -self.max_speed, self.max_speed)
        position += velocity
        position = np.clip(position, self.min_position, self.max_position)
        if position == self.min_position and velocity < 0:
            velocity = 0
        terminated = bool(position >= self.goal_position and velocity >= self.goal_velocity)
        reward = -1.0
        self.state = (position, velocity)
        return self._obs(), reward, terminated


class MountainCarContinuous(_BaseEnv):
    """gymnasium/envs/classic_control/continuous_mountain_car.py::Continuous_MountainCarEnv (restated from memory, like the others).

    What makes this one delicate is typing, not physics: reset() leaves `self.state` a FLOAT64 array ([uniform draw, 0]), step()
    leaves it a FLOAT32 array, the action arrives as a float32 array and the constants are python floats - so under NEP 50 the
    first step of an episode computes in float64 and every later one in float32, `0.0025 * math.cos(...)` is a python float that is
    rounded to float32 before the subtraction, and a clamped velocity / position / force is a python float again.  Writing the
    arithmetic with the same objects gymnasium uses reproduces all of that by construction."""

    env_id = "MountainCarContinuous-v0"
    default_max_episode_steps = 999

    def __init__(self, max_episode_steps=None, goal_velocity=0):
        super().__init__(max_episode_steps)
        self.min_action = -1.0
        self.max_action = 1.0
        self.min_position = -1.2
        self.max_position = 0.6
        self.max_speed = 0.07
        self.goal_position = 0.45
        self.goal_velocity = goal_velocity
        self.power = 0.0015
        self.observation_space = _Space(shape=(2,), dtype=np.float32)
        self.action_space = _Space(shape=(1,), dtype=np.float32, low=-1.0, high=1.0)

    def _draw_state(self):
        return np.array([self.np_random.uniform(low=-0.6, high=-0.4), 0])

    def _obs(self):
        return np.array(self.state, dtype=np.float32)

    def _physics(self, action):
        position = self.state[0]
        velocity = self.state[1]
        force = min(max(action[0], self.min_action), self.max_action)
        velocity += force * self.power - 0.0025 * math.cos(3 * position)
        if velocity > self.max_speed:
            velocity = self.max_speed
        if velocity < -self.max_speed:
            velocity = -self.max_speed
        position += velocity
        if position > self.max_position:
            position = self.max_position
        if position < self.min_position:
            position = self.min_position
        if position == self.min_position and velocity < 0:
            velocity = 0
        terminated = bool(position >= self.goal_position and velocity >= self.goal_velocity)
        reward = 0
        if terminated:
            reward = 100.0
        reward -= math.pow(action[0], 2) * 0.1
        self.state = np.array([position, velocity], dtype=np.float32)
        return self.state, reward, terminated


ENV_IDS = {"CartPole-v1": CartPole, "Pendulum-v1": Pendulum, "Acrobot-v1": Acrobot, "MountainCar-v0": MountainCar,
           "MountainCarContinuous-v0": MountainCarContinuous}


def make(env_id: str, max_episode_steps: int | None = None):
    return ENV_IDS[env_id](max_episode_steps=max_episode_steps)
