"""Drop-in `AsyncTools.utils` (reference: /root/reference/AsyncTools/utils.py:1-50): the seven free functions of the
rollout bookkeeping, same names / arguments / return conventions (host numpy in, host numpy out).

Each array computation runs in a CUDA kernel of libprl_b200.so - host arrays are uploaded, the kernel runs, the result
is downloaded; there is no numpy implementation behind them.  These per-step entry points exist for callers that write
their own loop (README "custom loop"); `AsyncPPO.worker()` does not come through here, it runs the whole episode loop
as one fused launch with the mask, the compaction and the buffer append inside the kernel.
"""
from __future__ import annotations

import numpy as np
import torch as t

from prl_b200 import ops
from prl_b200._lib import require_cuda


def _mask_dev(is_env_terminal):
    return t.from_numpy(np.ascontiguousarray(np.asarray(is_env_terminal, dtype=np.bool_)).view(np.uint8)).to(require_cuda())


def _active(is_env_terminal):
    """(idx int32 device tensor, n) of the non-terminal envs, ascending (stream compaction kernel)."""
    m = _mask_dev(is_env_terminal)
    idx, cnt = ops.compact_indices(m, want=False)
    return idx, int(cnt.item())


def indexes_of_active_environments(num_envs: int, is_env_terminal: np.ndarray):
    """utils.py:3-4: np.arange(num_envs)[~is_env_terminal]"""
    idx, n = _active(np.asarray(is_env_terminal)[:num_envs])
    return idx[:n].cpu().numpy().astype(np.int64)


def number_of_active_environments(is_env_terminal: np.ndarray):
    """utils.py:6-7: np.sum(~is_env_terminal)"""
    return np.int64(_active(is_env_terminal)[1])


def range_of_active_environments(is_env_terminal: np.ndarray):
    """utils.py:9-12: np.arange(number of active envs) - the compact ranks"""
    return np.arange(number_of_active_environments(is_env_terminal))


def inactive_states_dropout(states: np.ndarray, dones: np.ndarray):
    """utils.py:14-15: states[~dones] - stream compaction of the next observations (any float dtype, any row shape)."""
    states = np.asarray(states)
    keep, kc = ops.compact_indices(_mask_dev(dones), want=False)
    k = int(kc.item())
    if states.dtype == np.float32:
        rows = t.from_numpy(np.ascontiguousarray(states.reshape(len(states), -1))).to(require_cuda())
        out = ops.gather_rows(rows, keep, kc, len(states))[:k].cpu().numpy()
        return out.reshape((k,) + states.shape[1:])
    # other dtypes (float64 in the reference's unit test): move the raw bytes, 4-byte words, through the same kernel
    raw = np.ascontiguousarray(states).reshape(len(states), -1).view(np.float32)
    rows = t.from_numpy(raw).to(require_cuda())
    out = ops.gather_rows(rows, keep, kc, len(states))[:k].cpu().numpy()
    return out.view(states.dtype).reshape((k,) + states.shape[1:])


def buffer_append(buffer, states: np.ndarray, actions: np.ndarray, rewards: np.ndarray, dones: np.ndarray,
                  is_env_terminal: np.ndarray, num_envs: int):
    """utils.py:17-36: store step data of the active envs (compact rank i -> env index idx[i]) as float32 items."""
    if hasattr(buffer, "append_step"):  # our VecMemory: device append kernel
        buffer.append_step(states, actions, rewards, dones, is_env_terminal, num_envs)
        return
    idxs = indexes_of_active_environments(num_envs, is_env_terminal)  # duck-typed foreign buffer: its own push()
    for i_, idx_ in enumerate(idxs):
        buffer.push(idx_, states[i_], actions[i_], rewards[i_], dones[i_])


def update_active_environments_list(is_env_terminal: np.ndarray, dones: np.ndarray):
    """utils.py:38-43: is_env_terminal[active] = dones, in place, and returns it."""
    m = _mask_dev(is_env_terminal)
    idx, cnt = ops.compact_indices(m, want=False)
    n = int(cnt.item())
    dones = np.asarray(dones, dtype=np.bool_).reshape(-1)
    if len(dones) != n:
        raise ValueError(f"shape mismatch: {len(dones)} dones for {n} active environments")
    ops.mask_update(m, idx, _mask_dev(dones), n)
    is_env_terminal[...] = m.cpu().numpy().view(np.bool_)
    return is_env_terminal


def buffer_to_target_buffer_transfer(buffer, target_buffer):
    """utils.py:45-50: env-major, time-minor concatenation of the per-env episodes behind what the target already
    holds, then buffer.clear()."""
    if hasattr(buffer, "transfer_to") and buffer.transfer_to(target_buffer):
        return
    # host lists (assigned by the caller, or a foreign buffer / target): plain list concatenation of Python objects
    for name in ("states", "actions", "rewards", "dones"):
        dst = getattr(target_buffer, name)
        for per_env in getattr(buffer, name):
            dst += list(per_env)
    buffer.clear()
