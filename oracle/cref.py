"""ORACLE (test infrastructure): ctypes binding of oracle/c/prl_oracle.c.

`build()` compiles the C restatement with the recipe in oracle/Makefile; `lib()` loads it."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

ENV_CODES = {"CartPole-v1": 0, "Pendulum-v1": 1, "Acrobot-v1": 2, "MountainCar-v0": 3, "MountainCarContinuous-v0": 4}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "c", "prl_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "_build/liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_rollout.restype = C.c_int64
        L.orc_env_step.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def env_dims(env_id: str):
    v = [C.c_int() for _ in range(5)]
    assert lib().orc_env_dims(ENV_CODES[env_id], *[C.byref(x) for x in v]) == 0
    S, O, A, cont, ms = [x.value for x in v]
    return dict(S=S, O=O, A=A, continuous=bool(cont), max_steps=ms)


def env_step(env_id: str, state: np.ndarray, action):
    """Step ONE env in place. state: float64[S]. Returns (obs f32[O], reward f64, terminated)."""
    d = env_dims(env_id)
    obs = np.empty(d["O"], np.float32)
    r = C.c_double()
    if d["continuous"]:
        a = np.ascontiguousarray(action, np.float32)
    else:
        a = np.array([int(action)], np.int32)
    assert state.dtype == np.float64 and state.flags.c_contiguous
    term = lib().orc_env_step(ENV_CODES[env_id], _p(state), _p(a), _p(obs), C.byref(r))
    return obs, r.value, bool(term)


def env_obs(env_id: str, state: np.ndarray):
    d = env_dims(env_id)
    obs = np.empty(d["O"], np.float32)
    lib().orc_env_obs(ENV_CODES[env_id], _p(np.ascontiguousarray(state, np.float64)), _p(obs))
    return obs


def rollout(env_id: str, init_state: np.ndarray, tape: np.ndarray, max_steps: int):
    """Teacher-forced worker(): init_state f64[E,S]; tape int32[T,E] or f32[T,E,A]."""
    d = env_dims(env_id)
    E = init_state.shape[0]
    AS = d["A"] if d["continuous"] else 1
    cap = E * max_steps
    init_state = np.ascontiguousarray(init_state, np.float64)
    width = init_state.shape[1]
    if width < d["S"]:   # MountainCarContinuous: {position, velocity} + the "stepped" flag the C statement keeps (0 after reset)
        init_state = np.ascontiguousarray(np.concatenate([init_state, np.zeros((E, d["S"] - width))], 1))
    tape = np.ascontiguousarray(tape, np.float32 if d["continuous"] else np.int32)
    assert tape.shape[0] >= max_steps and tape.shape[1] == E
    fs = np.empty((cap, d["O"]), np.float32)
    fa = np.empty((cap, AS), np.float32)
    fr = np.empty(cap, np.float32)
    fd = np.empty(cap, np.float32)
    lens = np.empty(E, np.int32)
    fin = np.empty_like(init_state)
    rs = C.c_double()
    n = lib().orc_rollout(ENV_CODES[env_id], E, max_steps, _p(init_state), _p(tape), _p(fs), _p(fa), _p(fr),
                          _p(fd), _p(lens), _p(fin), C.byref(rs))
    assert n >= 0
    fa = fa[:n] if d["continuous"] else fa[:n, 0]
    return dict(states=fs[:n], actions=fa, rewards=fr[:n], dones=fd[:n], lengths=lens, final_state=fin[:, :width],
                reward_sum=rs.value, N=int(n))


def gae(rewards, dones, values, next_value, gamma, gae_lambda):
    r = np.ascontiguousarray(rewards, np.float32)
    d = np.ascontiguousarray(dones, np.float32)
    v = np.ascontiguousarray(values, np.float32)
    out = np.empty_like(r)
    lib().orc_gae(_p(r), _p(d), _p(v), C.c_float(float(next_value)), C.c_double(gamma), C.c_double(gae_lambda),
                  C.c_int64(r.size), _p(out))
    return out


def adv_norm(returns, values):
    r = np.ascontiguousarray(returns, np.float32)
    v = np.ascontiguousarray(values, np.float32)
    out = np.empty_like(r)
    m, s = C.c_double(), C.c_double()
    lib().orc_adv_norm(_p(r), _p(v), C.c_int64(r.size), _p(out), C.byref(m), C.byref(s))
    return out, m.value, s.value


def pow2(x64=None, x32=None):
    """libm pow(x, 2.0) / powf(x, 2.0f) elementwise - numpy's float64 / float32 scalar `x ** 2`."""
    n = len(x64) if x64 is not None else len(x32)
    o64 = np.empty(n, np.float64) if x64 is not None else None
    o32 = np.empty(n, np.float32) if x32 is not None else None
    lib().orc_pow_scalar(_p(np.ascontiguousarray(x64, np.float64)) if x64 is not None else None, C.c_double(2.0),
                         _p(o64) if o64 is not None else None,
                         _p(np.ascontiguousarray(x32, np.float32)) if x32 is not None else None, C.c_float(2.0),
                         _p(o32) if o32 is not None else None, C.c_int64(n))
    return o64, o32
