"""Thin torch-tensor front end of the C ABI: one Python function per entry point of include/prl_b200.h.

torch is plumbing only (device memory + streams); every function here launches hand-written sm_100a kernels
through ctypes and raises if the library is missing or a call fails.  All tensors must live on the current
CUDA device and be contiguous.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ACT_F32, ACT_I32, ACT_I64, call

NULL = None


def _ptr(t, dtype=None):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "prl_b200 ops need contiguous CUDA tensors"
    if dtype is not None:
        assert t.dtype == dtype, f"expected {dtype}, got {t.dtype}"
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


def policy_param_count(is_continuous, O, A):
    return int(_lib.fn("prl_policy_param_count")(int(is_continuous), int(O), int(A)))


def rnd_param_count(I, Oo):
    return int(_lib.fn("prl_rnd_param_count")(int(I), int(Oo)))


# ------------------------------------------------------------------------------------------------ test hooks
def test_sincos(x):
    s, c = torch.empty_like(x), torch.empty_like(x)
    call("prl_test_sincos", _ptr(x, torch.float64), _ptr(s), _ptr(c), x.numel(), _stream())
    return s, c


def test_pow2(x64=None, x32=None):
    n = (x64 if x64 is not None else x32).numel()
    o64 = torch.empty_like(x64) if x64 is not None else None
    o32 = torch.empty_like(x32) if x32 is not None else None
    call("prl_test_pow2", _ptr(x64), _ptr(o64), _ptr(x32), _ptr(o32), n, _stream())
    return o64, o32


def test_philox(seed, c0, c1, c2, c3):
    out = torch.empty(4, dtype=torch.int32, device=_dev())
    call("prl_test_philox", seed, c0, c1, c2, c3, _ptr(out), _stream())
    return out.cpu().numpy().view("uint32")


def test_pcg64(seeds, draws):
    """[n][draws] uint64 outputs of np.random.PCG64(np.random.SeedSequence(seed)) computed on the device."""
    import numpy as np
    sd = torch.from_numpy(np.asarray(seeds, dtype=np.uint64).view(np.int64)).to(_dev())
    out = torch.empty(len(sd), draws, dtype=torch.int64, device=_dev())
    call("prl_test_pcg64", _ptr(sd), len(sd), draws, _ptr(out), _stream())
    return out.cpu().numpy().view(np.uint64)


def test_umma(mode, A, B, cfg=None):
    n_out = {0: 128, 1: 64, 2: 64, 3: 16}[mode] if cfg is None else cfg[2]
    D = torch.full((128, n_out), float("nan"), dtype=torch.float32, device=_dev())
    status = torch.zeros(2, dtype=torch.int32, device=_dev())   # [0] = timed out, [1] = cycles from first issue to completion
    if cfg is not None:
        cfg = list(cfg) + [0, 0, 1][len(cfg) - 13:] if len(cfg) < 16 else list(cfg)
        cfg = cfg + [0] * (17 - len(cfg))   # [16]: lane offset of the accumulator address
    cfg_arr = (C.c_int32 * 17)(*[int(x) for x in cfg]) if cfg is not None else None
    call("prl_test_umma", -1 if cfg is not None else mode, _ptr(A, torch.float32), _ptr(B, torch.float32), _ptr(D), _ptr(status),
         cfg_arr, _stream())
    st = status.cpu().numpy()
    test_umma.last_cycles = int(st[1])
    return D, int(st[0])


# ------------------------------------------------------------------------------------------------ env state
class EnvState:
    """Device-resident state of E copies of one classic-control env: SoA fp64 [S][E], TimeLimit counters,
    terminal mask (True = terminal, the reference's `envs_active`)."""

    def __init__(self, env_id: str, num_envs: int, max_episode_steps: int | None = None):
        self.info = _lib.env_info(env_id)
        self.env_id, self.code, self.E = env_id, self.info["code"], int(num_envs)
        self.max_steps = int(max_episode_steps or self.info["max_steps"])
        d = _dev()
        self.state = torch.zeros(self.info["S"], self.E, dtype=torch.float64, device=d)
        self.elapsed = torch.zeros(self.E, dtype=torch.int32, device=d)
        self.terminal = torch.zeros(self.E, dtype=torch.uint8, device=d)

    def reset(self, seed: int, episode: int):
        obs = torch.empty(self.E, self.info["O"], dtype=torch.float32, device=_dev())
        call("prl_env_reset", self.code, self.E, seed, episode, _ptr(self.state), _ptr(self.elapsed),
             _ptr(self.terminal), _ptr(obs), _stream())
        return obs

    def reset_numpy(self, seeds=None):
        """reset() from numpy's seeded stream (csrc/np_rng.cuh).  seeds: uint64-valued int64 CUDA tensor [E] -> every env's
        generator is re-created as PCG64(SeedSequence(seed)); None -> the stored generators continue."""
        if seeds is None and getattr(self, "np_rng", None) is None:
            raise ValueError("reset_numpy(): the generators were never seeded")
        if seeds is not None:
            assert tuple(seeds.shape) == (self.E,)
            if getattr(self, "np_rng", None) is None:
                self.np_rng = torch.zeros(4, self.E, dtype=torch.int64, device=_dev())   # {state_hi, state_lo, inc_hi, inc_lo}
        obs = torch.empty(self.E, self.info["O"], dtype=torch.float32, device=_dev())
        call("prl_env_reset_numpy", self.code, self.E, _ptr(seeds, torch.int64) if seeds is not None else None,
             _ptr(self.np_rng), _ptr(self.state), _ptr(self.elapsed), _ptr(self.terminal), _ptr(obs), _stream())
        return obs

    def set_state(self, state_aos):
        """state_aos: [E][S] float64 CUDA tensor (injected start states)."""
        assert tuple(state_aos.shape) == (self.E, self.info["S"])
        obs = torch.empty(self.E, self.info["O"], dtype=torch.float32, device=_dev())
        call("prl_env_set_state", self.code, self.E, _ptr(state_aos, torch.float64), _ptr(self.state),
             _ptr(self.elapsed), _ptr(self.terminal), _ptr(obs), _stream())
        return obs

    def get_state(self):
        out = torch.empty(self.E, self.info["S"], dtype=torch.float64, device=_dev())
        call("prl_env_get_state", self.code, self.E, _ptr(self.state), _ptr(out), _stream())
        return out

    def step(self, active_idx, n, actions):
        """Compact step of the n active envs. actions: int32/int64 [n] or float32 [n][A]."""
        O = self.info["O"]
        d = _dev()
        obs = torch.empty(n, O, dtype=torch.float32, device=d)
        rewards = torch.empty(n, dtype=torch.float64, device=d)
        dones = torch.empty(n, dtype=torch.uint8, device=d)
        truncs = torch.empty(n, dtype=torch.uint8, device=d)
        adt = {torch.int32: ACT_I32, torch.int64: ACT_I64, torch.float32: ACT_F32}[actions.dtype]
        call("prl_env_step", self.code, self.E, n, _ptr(active_idx, torch.int32), _ptr(actions), adt, _ptr(self.state),
             _ptr(self.elapsed), self.max_steps, _ptr(obs), _ptr(rewards), _ptr(dones), _ptr(truncs), _stream())
        return obs, rewards, dones, truncs


# ------------------------------------------------------------------------------------------------ utils kernels
def _ws(nbytes):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=_dev())


def compact_indices(flags, want: bool):
    """idx = arange(n)[flags == want] (ascending), count as a device int32 scalar tensor."""
    n = flags.numel()
    idx = torch.empty(n, dtype=torch.int32, device=_dev())
    count = torch.zeros(1, dtype=torch.int32, device=_dev())
    ws = _ws(_lib.fn("prl_scan_ws_bytes")(n))
    call("prl_compact_indices", _ptr(flags, torch.uint8), n, int(want), _ptr(idx), _ptr(count), _ptr(ws), ws.numel(), _stream())
    return idx, count


def gather_rows(rows, idx, count, max_rows):
    width = rows.shape[1] if rows.dim() > 1 else 1
    out = torch.empty((max_rows,) + tuple(rows.shape[1:]), dtype=torch.float32, device=_dev())
    call("prl_gather_rows", _ptr(rows, torch.float32), _ptr(idx, torch.int32), _ptr(count, torch.int32), max_rows, width,
         _ptr(out), _stream())
    return out


def mask_update(terminal, active_idx, dones, n):
    call("prl_mask_update", _ptr(terminal, torch.uint8), _ptr(active_idx, torch.int32), _ptr(dones, torch.uint8), n, _stream())


class RolloutBuffer:
    """Device VecMemory: time-major [T][C][E] float32 slots + per-env lengths (AsyncPPO.py:11-33)."""

    def __init__(self, num_envs, T_cap, obs_dim, act_width):
        d = _dev()
        self.E, self.T, self.O, self.AW = int(num_envs), int(T_cap), int(obs_dim), int(act_width)
        self.states = torch.empty(self.T, self.O, self.E, dtype=torch.float32, device=d)
        self.actions = torch.empty(self.T, self.AW, self.E, dtype=torch.float32, device=d)
        self.rewards = torch.empty(self.T, self.E, dtype=torch.float32, device=d)
        self.dones = torch.empty(self.T, self.E, dtype=torch.float32, device=d)
        self.lengths = torch.zeros(self.E, dtype=torch.int32, device=d)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=d)
        self._ws = _ws(_lib.fn("prl_scan_ws_bytes")(self.E))
        # what the fused worker can produce besides the reference's four fields (allocated by want_eval()): the acting policy's
        # log-prob and state value of every transition, and the GAE returns computed on the time-major planes
        self.logp = self.values = self.returns = None

    def want_eval(self):
        if self.logp is None:
            d = _dev()
            self.logp, self.values, self.returns = (torch.empty(self.T, self.E, dtype=torch.float32, device=d) for _ in range(3))
        return self

    def append(self, active_idx, n, states, actions, rewards, dones):
        call("prl_buffer_append", self.E, self.T, n, _ptr(active_idx, torch.int32), _ptr(states, torch.float32), self.O,
             _ptr(actions, torch.float32), self.AW, _ptr(rewards, torch.float32), _ptr(dones, torch.float32),
             _ptr(self.states), _ptr(self.actions), _ptr(self.rewards), _ptr(self.dones), _ptr(self.lengths),
             _ptr(self.overflow), _stream())

    def transfer(self, mem_states, mem_actions, mem_rewards, mem_dones, base, total_out, extra=()):
        """`extra`: (plane [T][E], rows [cap]) pairs that travel with the four fields (prl_buffer_transfer_ex)."""
        cap = mem_rewards.numel()
        n = len(extra)
        planes = (C.c_void_p * max(n, 1))(*[p.data_ptr() for p, _ in extra])
        rows = (C.c_void_p * max(n, 1))(*[r.data_ptr() for _, r in extra])
        for p, r in extra:
            assert p.is_cuda and p.is_contiguous() and p.dtype == torch.float32 and tuple(p.shape) == (self.T, self.E)
            assert r.is_cuda and r.is_contiguous() and r.dtype == torch.float32 and r.numel() >= cap
        call("prl_buffer_transfer_ex", self.E, self.T, self.O, self.AW, _ptr(self.states), _ptr(self.actions),
             _ptr(self.rewards), _ptr(self.dones), n, planes if n else None, rows if n else None, _ptr(self.lengths), base, cap,
             _ptr(mem_states, torch.float32), _ptr(mem_actions, torch.float32), _ptr(mem_rewards, torch.float32),
             _ptr(mem_dones, torch.float32), _ptr(total_out, torch.int64), _ptr(self._ws), self._ws.numel(), _stream())


# ------------------------------------------------------------------------------------------------ policy
def policy_act(params, is_continuous, O, A, action_scaling, states, seed, call_index, row_ids=None, want_dist=False):
    n = states.shape[0]
    d = _dev()
    actions = torch.empty((n, A), dtype=torch.float32, device=d) if is_continuous else torch.empty(n, dtype=torch.int64, device=d)
    dist = torch.empty((n, 2 * A if is_continuous else A), dtype=torch.float32, device=d) if want_dist else None
    call("prl_policy_act", _ptr(params, torch.float32), int(is_continuous), O, A, float(action_scaling or 1.0),
         _ptr(states, torch.float32), _ptr(row_ids), n, seed, call_index, _ptr(actions), _ptr(dist), _stream())
    return (actions, dist) if want_dist else actions


def policy_evaluate(params, is_continuous, O, A, states, actions, entropy_sum=None, logp=None, value=None):
    n = states.shape[0]
    d = _dev()
    logp = torch.empty(n, dtype=torch.float32, device=d) if logp is None else logp
    value = torch.empty(n, dtype=torch.float32, device=d) if value is None else value
    if entropy_sum is None:
        entropy_sum = torch.zeros(1, dtype=torch.float64, device=d)
    call("prl_policy_evaluate", _ptr(params, torch.float32), int(is_continuous), O, A, _ptr(states, torch.float32),
         _ptr(actions, torch.float32), n, _ptr(logp), _ptr(value), _ptr(entropy_sum, torch.float64), _stream())
    return logp, value, entropy_sum


def rollout(env: EnvState, buf: RolloutBuffer, params, action_scaling, seed, episode, scores, tape=None, evaluate=False, auto_reset_horizon=0, steps=None):
    """Fused AsyncPPO.worker(): one launch.  scores: float64[2] device tensor, accumulated.  evaluate: also fill buf.logp / buf.values
    (the old-policy evaluation of PPO.learn, bit-identical to policy_evaluate on the same rows).  auto_reset_horizon > 0: opt-in
    auto-reset (every env fills all buf.T slots; an episode ends at termination or after that many steps)."""
    if evaluate:
        buf.want_eval()
    T = buf.T if steps is None else int(steps)
    assert 0 < T <= buf.T
    if getattr(env, "score_ws", None) is None:   # fixed-order accumulation of the reward sum (bit-reproducible scores)
        env.score_ws = torch.zeros(int(_lib.fn("prl_rollout_score_ws_doubles")(env.E)), dtype=torch.float64, device=_dev())
    call("prl_rollout_eval", env.code, env.E, T, _ptr(params), float(action_scaling or 1.0), seed, episode, _ptr(tape),
         _ptr(env.state), _ptr(env.elapsed), _ptr(env.terminal), _ptr(buf.states), _ptr(buf.actions), _ptr(buf.rewards),
         _ptr(buf.dones), _ptr(buf.logp) if evaluate else None, _ptr(buf.values) if evaluate else None, _ptr(buf.lengths),
         _ptr(scores, torch.float64), int(auto_reset_horizon), _ptr(env.score_ws, torch.float64), _stream())


# ------------------------------------------------------------------------------------------------ GAE
def gae(rewards, dones, values, gamma, gae_lambda, next_value=None, out=None, ws=None):
    N = rewards.numel()
    out = torch.empty(N, dtype=torch.float32, device=_dev()) if out is None else out
    need = _lib.fn("prl_gae_ws_bytes")(N)
    ws = _ws(need) if ws is None or ws.numel() < need else ws
    call("prl_gae", _ptr(rewards, torch.float32), _ptr(dones, torch.float32), _ptr(values, torch.float32),
         _ptr(next_value), float(gamma), float(gae_lambda), N, _ptr(out), _ptr(ws), ws.numel(), _stream())
    return out


def gae_columns(rewards, dones, values, lengths, gamma, gae_lambda, out=None):
    T, E = rewards.shape
    out = torch.empty_like(rewards) if out is None else out
    call("prl_gae_columns", _ptr(rewards, torch.float32), _ptr(dones, torch.float32), _ptr(values, torch.float32),
         _ptr(lengths), E, T, float(gamma), float(gae_lambda), _ptr(out), _stream())
    return out


def adv_normalize(returns, values, stats=None, phase=3, out=None):
    N = returns.numel()
    d = _dev()
    if stats is None:
        stats = torch.zeros(4, dtype=torch.float64, device=d)
    if out is None and (phase & 2):
        out = torch.empty(N, dtype=torch.float32, device=d)
    call("prl_adv_normalize", _ptr(returns, torch.float32), _ptr(values, torch.float32), N, _ptr(out), _ptr(stats, torch.float64),
         phase, _stream())
    return out, stats


# ------------------------------------------------------------------------------------------------ update
def update_ws_floats(is_continuous, O, A, batch):
    return int(_lib.fn("prl_update_ws_floats")(int(is_continuous), O, A, batch))


def ppo_grad(params, is_continuous, O, A, states, actions, old_logp, adv, returns, policy_clip, inv_count, grad, loss_out, ws):
    b = states.shape[0]
    call("prl_ppo_grad", _ptr(params, torch.float32), int(is_continuous), O, A, _ptr(states, torch.float32),
         _ptr(actions, torch.float32), _ptr(old_logp, torch.float32), _ptr(adv, torch.float32), _ptr(returns, torch.float32),
         b, float(policy_clip), float(inv_count), _ptr(grad, torch.float32), _ptr(loss_out, torch.float64),
         _ptr(ws, torch.float32), ws.numel(), _stream())


def tc_supported(is_continuous, O, A):
    """0: no tensor-core update for this policy; 1: every form (gradient, fused step, sharded fused step); 2: the gradient form only
    (continuous policies: pre-pass + two passes of the two-head kernel inside prl_ppo_grad_tc)."""
    return int(_lib.fn("prl_ppo_grad_tc_supported")(int(is_continuous), int(O), int(A)))


def update_tc_ws_floats(is_continuous, O, A, batch):
    return int(_lib.fn("prl_update_tc_ws_floats")(int(is_continuous), O, A, batch))


def ppo_grad_tc(params, is_continuous, O, A, states, actions, old_logp, adv, returns, policy_clip, inv_count, grad, loss_out, ws):
    b = states.shape[0]
    call("prl_ppo_grad_tc", _ptr(params, torch.float32), int(is_continuous), O, A, _ptr(states, torch.float32),
         _ptr(actions, torch.float32), _ptr(old_logp, torch.float32), _ptr(adv, torch.float32), _ptr(returns, torch.float32),
         b, float(policy_clip), float(inv_count), _ptr(grad, torch.float32), _ptr(loss_out, torch.float64),
         _ptr(ws, torch.float32), ws.numel(), _stream())


def ppo_step_tc(params, is_continuous, O, A, states, actions, old_logp, adv, returns, policy_clip, inv_count, grad, loss_out, opt, ws):
    """Gradient + clip_grad_norm_ + AdamW in one cooperative launch; `opt` is a prl_b200.optim.FusedAdamW."""
    b = states.shape[0]
    call("prl_ppo_step_tc", _ptr(params, torch.float32), int(is_continuous), O, A, _ptr(states, torch.float32),
         _ptr(actions, torch.float32), _ptr(old_logp, torch.float32), _ptr(adv, torch.float32), _ptr(returns, torch.float32),
         b, float(policy_clip), float(inv_count), _ptr(grad, torch.float32), _ptr(loss_out, torch.float64),
         _ptr(opt.exp_avg, torch.float32), _ptr(opt.exp_avg_sq, torch.float32), _ptr(opt.step_dev, torch.int64), float(opt.lr),
         float(opt.weight_decay), float(opt.max_norm), _ptr(opt.grad_norm, torch.float64), _ptr(ws, torch.float32), ws.numel(), _stream())


def ppo_step_tc_p2p(params, is_continuous, O, A, states, actions, old_logp, adv, returns, policy_clip, inv_count, grad, loss_out, opt, ws, xch):
    """Sharded prl_ppo_step_tc: gradient exchange over peer memory inside the kernel (`xch`: prl_b200.dist.PeerExchange)."""
    b = states.shape[0] if states is not None else 0
    call("prl_ppo_step_tc_p2p", _ptr(params, torch.float32), int(is_continuous), O, A, _ptr(states), _ptr(actions), _ptr(old_logp),
         _ptr(adv), _ptr(returns), b, float(policy_clip), float(inv_count), _ptr(grad, torch.float32), _ptr(loss_out, torch.float64),
         _ptr(opt.exp_avg, torch.float32), _ptr(opt.exp_avg_sq, torch.float32), _ptr(opt.step_dev, torch.int64), float(opt.lr),
         float(opt.weight_decay), float(opt.max_norm), _ptr(opt.grad_norm, torch.float64), _ptr(xch.table, torch.int64), xch.rank,
         xch.world_size, _ptr(ws, torch.float32), ws.numel(), _stream())


def ppo_grad_tc_status(ws):
    st = C.c_int(0)
    call("prl_ppo_grad_tc_status", _ptr(ws, torch.float32), C.byref(st), _stream())
    return st.value


def adamw_step(params, grad, exp_avg, exp_avg_sq, step, lr, weight_decay=0.01, max_norm=2.0, grad_norm_out=None):
    call("prl_adamw_step", _ptr(params, torch.float32), _ptr(grad, torch.float32), _ptr(exp_avg, torch.float32),
         _ptr(exp_avg_sq, torch.float32), params.numel(), int(step), float(lr), float(weight_decay), float(max_norm),
         _ptr(grad_norm_out), _stream())


def adamw_step_dev(params, grad, exp_avg, exp_avg_sq, step_counter, lr, weight_decay=0.01, max_norm=2.0, grad_norm_out=None):
    call("prl_adamw_step_dev", _ptr(params, torch.float32), _ptr(grad, torch.float32), _ptr(exp_avg, torch.float32),
         _ptr(exp_avg_sq, torch.float32), params.numel(), _ptr(step_counter, torch.int64), float(lr), float(weight_decay),
         float(max_norm), _ptr(grad_norm_out), _stream())


def rnd_intrinsic(target_params, pred_params, I, Oo, states, beta, add_to=None, out=None):
    n = states.shape[0]
    out = torch.empty(n, dtype=torch.float32, device=_dev()) if out is None else out
    call("prl_rnd_intrinsic", _ptr(target_params, torch.float32), _ptr(pred_params, torch.float32), I, Oo,
         _ptr(states, torch.float32), n, float(beta), _ptr(add_to), _ptr(out), _stream())
    return out


def rnd_grad(target_params, pred_params, I, Oo, states, grad, loss_out, ws):
    call("prl_rnd_grad", _ptr(target_params, torch.float32), _ptr(pred_params, torch.float32), I, Oo,
         _ptr(states, torch.float32), states.shape[0], _ptr(grad, torch.float32), _ptr(loss_out, torch.float64),
         _ptr(ws, torch.float32), ws.numel(), _stream())
