"""Drop-in `AsyncTools.AsyncPPO` (reference: /root/reference/AsyncTools/AsyncPPO.py:11-165): `VecMemory`,
`EnvVectorizer` and the rollout driver `AsyncPPO`, same constructors / methods / attributes.

B200 design: the `num_envs` environments are not Python objects stepped in a loop but SoA fp64 state in HBM
(prl_b200.ops.EnvState) stepped one env per thread; the per-env transition lists are a time-major [T][C][E] float32
buffer in HBM (ops.RolloutBuffer).  `AsyncPPO.worker()` - one episode per env, finished envs drop out, no auto-reset -
runs as ONE kernel launch (prl_rollout: policy forward, sampling, physics, TimeLimit, done|truncate mask and buffer
write fused, looped over time inside the kernel) followed by the env-major transfer into `ppo.memory`.
The per-step API (`EnvVectorizer.step`, `utils.*`) is kept for callers that write their own loop and goes through the
same physics / policy / compaction kernels one step at a time, with host numpy at the boundary like the reference.
"""
from __future__ import annotations

import numpy as np
import torch as t
from tqdm import tqdm

import AsyncTools.utils as utils
from prl_b200 import dist as pdist
from prl_b200 import ops
from prl_b200._lib import require_cuda
from prl_b200.envs import describe

_FIELDS = ("states", "actions", "rewards", "dones")


class VecMemory:
    """Per-env transition store (AsyncPPO.py:11-33).

    Two representations behind the reference's attributes:
      * device (what the rollout uses): ops.RolloutBuffer, time-major slots + per-env lengths, filled by the append
        kernel / the fused rollout; `.states` etc. then materialise a list-of-lists snapshot on demand;
      * host lists: what `push()` or direct assignment (`buffer.states[i] = [...]`) by user code produces - arbitrary
        Python objects, kept as they are.
    """

    def __init__(self, num_envs: int):
        self.num_envs = int(num_envs)
        self._host = {name: [[] for _ in range(self.num_envs)] for name in _FIELDS}
        self._host_used = False
        self.dev = None          # ops.RolloutBuffer, allocated on first device use
        self._dev_steps = 0      # appends since clear(): upper bound of every env's length
        self._scalar_actions = True

    # ---- reference surface ------------------------------------------------------------------------------------
    def _lists(self, name):
        if self._dev_steps and not self._host_used:
            return self._materialise()[name]
        self._host_used = True  # the caller may mutate / assign the lists it gets
        return self._host[name]

    states = property(lambda s: s._lists("states"))
    actions = property(lambda s: s._lists("actions"))
    rewards = property(lambda s: s._lists("rewards"))
    dones = property(lambda s: s._lists("dones"))

    def push(self, idx: int, state, action, reward, done):
        """AsyncPPO.py:20-24: every item is stored as a float32 copy."""
        self._host_used = True
        for name, x in zip(_FIELDS, (state, action, reward, done)):
            self._host[name][idx].append(np.asarray(x).astype(np.float32))

    def clear(self):
        for name in _FIELDS:
            for lst in self._host[name]:
                del lst[:]
        self._host_used = False
        if self.dev is not None and self._dev_steps:
            self.dev.lengths.zero_()
        self._dev_steps = 0

    # ---- device side ------------------------------------------------------------------------------------------
    def device(self, T_cap: int, obs_dim: int, act_width: int) -> "ops.RolloutBuffer":
        """The device buffer with room for `T_cap` steps per env (re-allocated, keeping contents, when it grows)."""
        d = self.dev
        if d is None or d.O != obs_dim or d.AW != act_width:
            self.dev = ops.RolloutBuffer(self.num_envs, T_cap, obs_dim, act_width)
        elif d.T < T_cap:
            new = ops.RolloutBuffer(self.num_envs, T_cap, obs_dim, act_width)
            for name in _FIELDS:
                getattr(new, name)[: d.T].copy_(getattr(d, name))
            new.lengths.copy_(d.lengths)
            self.dev = new
        return self.dev

    def append_step(self, states, actions, rewards, dones, is_env_terminal, num_envs):
        """utils.buffer_append on the device: rank i of the active set -> env idx[i], slot lengths[idx[i]]."""
        if self._host_used:  # the caller is building host lists: stay with them
            idxs = utils.indexes_of_active_environments(num_envs, is_env_terminal)
            for i_, idx_ in enumerate(idxs):
                self.push(idx_, states[i_], actions[i_], rewards[i_], dones[i_])
            return
        dev = require_cuda()
        f32 = lambda x, shape: t.from_numpy(np.ascontiguousarray(np.asarray(x).astype(np.float32).reshape(shape))).to(dev)  # noqa: E731
        idx, n = utils._active(np.asarray(is_env_terminal)[:num_envs])
        if n == 0:
            return
        s = f32(np.asarray(states)[:n], (n, -1))
        a_np = np.asarray(actions)[:n]
        self._scalar_actions = a_np.ndim == 1
        a = f32(a_np, (n, -1))
        cap = max(64, 1 << int(self._dev_steps + 1).bit_length())
        buf = self.device(max(cap, self.dev.T if self.dev is not None else 0), s.shape[1], a.shape[1])
        buf.append(idx, n, s, a, f32(np.asarray(rewards)[:n], (n,)), f32(np.asarray(dones)[:n], (n,)))
        self._dev_steps += 1

    def transfer_to(self, target) -> bool:
        """utils.buffer_to_target_buffer_transfer on the device; False when either side holds host lists."""
        if self._host_used or not hasattr(target, "reserve"):
            if self._dev_steps:  # device rows going to a foreign target: hand over host copies
                snap = self._materialise()
                for name in _FIELDS:
                    dst = getattr(target, name)
                    for per_env in snap[name]:
                        dst += per_env
                self.clear()
                return True
            return False
        if not self._dev_steps:
            self.clear()
            return True
        buf = self.dev
        n_new = int(buf.lengths.sum().item())
        target.append_from_rollout(buf, n_new, self._scalar_actions)
        self._dev_steps = 0
        return True

    def _materialise(self):
        buf = self.dev
        lens = buf.lengths.cpu().numpy()
        T = int(lens.max()) if len(lens) else 0
        out = {}
        for name in _FIELDS:
            arr = getattr(buf, name)[:T].cpu().numpy()  # [T, C, E] or [T, E]
            per_env = []
            for e in range(self.num_envs):
                col = arr[: lens[e], ..., e]
                if name == "actions" and self._scalar_actions:
                    col = col[:, 0]
                per_env.append(list(col))
            out[name] = per_env
        return out


class EnvVectorizer:
    """`num_envs` copies of one classic-control env, stepped by CUDA kernels (AsyncPPO.py:35-102).

    `env`: a `prl_b200.make("CartPole-v1")` descriptor, or any gymnasium env exposing `.spec.id` for which a kernel
    exists.  `envs_active` keeps the reference's meaning (True = that env's episode is over) and may be read and
    assigned by the caller, as the reference's loop does.
    """

    def __init__(self, env, num_envs: int = 1):
        self.device = require_cuda()
        self.desc = describe(env)
        self.num_envs = int(num_envs)
        self.envs = [env] * self.num_envs  # the reference keeps num_envs deep copies; here they share one descriptor
        self.action_space = env.action_space
        self.observation_space = env.observation_space
        self.sim = ops.EnvState(self.desc.env_id, self.num_envs, self.desc.max_episode_steps)
        self.seed = int(t.randint(0, 2 ** 62, (1,)).item())  # reset stream of the Philox generator
        if pdist.active() is not None:   # env-sharded run: identically seeded ranks must not reset their shards identically
            self.seed = pdist.rank_seed(self.seed, pdist.active().rank)
        self.episode = 0   # incremented by every reset(); selects the reset / action random streams
        self.t = 0         # steps since reset()

    # True = terminal, as in the reference (AsyncPPO.py:42)
    @property
    def envs_active(self) -> np.ndarray:
        return self.sim.terminal.cpu().numpy().view(np.bool_)

    @envs_active.setter
    def envs_active(self, value):
        v = np.ascontiguousarray(np.asarray(value, dtype=np.bool_).reshape(self.num_envs))
        self.sim.terminal.copy_(t.from_numpy(v.view(np.uint8)))

    def reset_device(self, seed=None) -> t.Tensor:
        self.episode += 1
        self.t = 0
        if seed is not None:
            # numpy-identical seeded resets: env i becomes what `envs[i].reset(seed=seeds[i])` gives in the reference
            # (gymnasium: Generator(PCG64(SeedSequence(seed))).uniform(low, high)); an int seeds env i with seed + i
            seeds = np.arange(self.num_envs, dtype=np.uint64) + np.uint64(seed) if np.isscalar(seed) else np.asarray(seed, dtype=np.uint64)
            if seeds.shape != (self.num_envs,):
                raise ValueError(f"seed must be an int or {self.num_envs} ints")
            self._numpy_stream = True
            return self.sim.reset_numpy(t.from_numpy(seeds.view(np.int64)).to(self.device))
        if getattr(self, "_numpy_stream", False):
            return self.sim.reset_numpy()   # every env's generator continues, as the reference's per-env generators do
        return self.sim.reset(self.seed, self.episode)

    def reset(self, seed=None):
        """AsyncPPO.py:48-62 -> (obs [E, O] float32, infos).  `seed` (an extension; the reference's reset() cannot be
        seeded): int or num_envs ints - see reset_device."""
        obs = self.reset_device(seed)
        return obs.cpu().numpy(), [{} for _ in range(self.num_envs)]

    def reset_to_device(self, states) -> t.Tensor:
        """Teacher forcing: start every env from the given fp64 state [E, S] (numpy, host tensor - pinned memory is
        copied asynchronously - or CUDA tensor) instead of a random draw.  Returns the observations on the device."""
        self.episode += 1
        self.t = 0
        if isinstance(states, np.ndarray):
            states = t.from_numpy(np.ascontiguousarray(states, dtype=np.float64))
        s = states.to(device=self.device, dtype=t.float64, non_blocking=True).contiguous()
        S = self.sim.info["S"]
        if s.dim() == 2 and s.shape[1] < S:   # trailing bookkeeping components (MountainCarContinuous' "stepped" flag) start at 0, as after reset()
            s = t.cat([s, t.zeros(s.shape[0], S - s.shape[1], dtype=t.float64, device=self.device)], 1).contiguous()
        return self.sim.set_state(s)

    def reset_to(self, states):
        return self.reset_to_device(states).cpu().numpy(), [{} for _ in range(self.num_envs)]

    def active_indices(self):
        idx, cnt = ops.compact_indices(self.sim.terminal, want=False)
        return idx, int(cnt.item())

    def step_device(self, idx, n, actions: t.Tensor):
        self.t += 1
        return self.sim.step(idx, n, actions)

    def step(self, actions: np.ndarray):
        """AsyncPPO.py:64-102: steps the non-terminal envs with the compact actions[rank]; compact results."""
        idx, n = self.active_indices()
        a = np.asarray(actions)
        if self.desc.is_continuous:
            a_dev = t.from_numpy(np.ascontiguousarray(a[:n], dtype=np.float32).reshape(n, -1)).to(self.device)
        else:
            a_dev = t.from_numpy(np.ascontiguousarray(a[:n].reshape(n), dtype=np.int64)).to(self.device)
        obs, rewards, dones, truncs = self.step_device(idx, n, a_dev)
        infos = np.array([{} for _ in range(n)], dtype=object)
        return (obs.cpu().numpy(), rewards.cpu().numpy(), dones.cpu().numpy().view(np.bool_),
                truncs.cpu().numpy().view(np.bool_), infos)

    def close(self):
        pass


class AsyncPPO:
    def __init__(self, env, ppo: object, num_envs: int = 32, steps: int = 100000):
        self.env = EnvVectorizer(env, num_envs)

        self.num_envs = num_envs
        self.steps = steps
        self.ppo = ppo

        self.step_score = np.array(0, dtype=np.int32)
        self.reward_score = np.array(0.0, dtype=np.float32)

        self.buffer = VecMemory(num_envs)
        self.fused = True           # set False to force the step-by-step loop (same results, one launch set per step)
        # opt-in, NOT the reference's worker (where a finished env drops out until the next worker(), AsyncPPO.py:118,143-146): with
        # auto_reset a finished env is reset inside the rollout kernel and keeps stepping, so every worker() yields exactly
        # num_envs x rollout_steps transitions (rollout_steps defaults to the env's TimeLimit; the last slot closes the running
        # episode with done = 1).  Fused worker only.
        self.auto_reset = False
        self.rollout_steps = None
        self.show_progress = True
        self._scores = t.zeros(2, dtype=t.float64, device=self.env.device)

    # ------------------------------------------------------------------------------------------------ rollout
    def _can_fuse(self) -> bool:
        from PPO.PPO import PPO as _PPO

        p = self.ppo
        d = self.env.desc
        return (self.fused and isinstance(p, _PPO) and type(p).get_action is _PPO.get_action and type(self.env) is EnvVectorizer
                and type(self.buffer) is VecMemory and not self.buffer._host_used and self.buffer._dev_steps == 0
                and p.observ_dim == d.observ_dim and p.action_dim == d.action_dim and bool(p.is_continuous) == d.is_continuous)

    def worker(self, initial_states=None):
        """AsyncPPO.py:117-146: one episode per env into ppo.memory.  `initial_states` (optional, beyond the reference's
        signature): host or device fp64 [E, S] start states to use instead of a random reset."""
        if self._can_fuse():
            self._worker_fused(initial_states)
        elif self.auto_reset:
            raise RuntimeError("auto_reset is an option of the fused worker (the reference's step-by-step loop has no auto-reset)")
        else:
            self._worker_stepwise(initial_states)

    def _worker_fused(self, initial_states=None):
        env, ppo, d = self.env, self.ppo, self.env.desc
        if initial_states is None:
            env.reset_device()
        else:
            env.reset_to_device(initial_states)
        aw = d.action_dim if d.is_continuous else 1
        T, horizon = env.sim.max_steps, 0
        if self.auto_reset:
            T, horizon = int(self.rollout_steps or env.sim.max_steps), env.sim.max_steps
        buf = self.buffer.device(T, d.observ_dim, aw)
        self._scores.zero_()
        # the old-policy evaluation of PPO.learn (PPO.py:134-154) rides along: the acting policy IS policy_old, so the log-prob of
        # the sampled action and V(s) are by-products of the step (bit-identical to the separate pass); without RND the rewards are
        # final too, so the GAE returns are computed right here on the time-major planes (coalesced across envs)
        fuse_eval = bool(getattr(ppo, "fuse_evaluation", False)) and (not d.is_continuous or d.action_dim == 1)
        ops.rollout(env.sim, buf, ppo.policy_old.flat, ppo._action_scale(), ppo._seed, env.episode, self._scores, evaluate=fuse_eval,
                    auto_reset_horizon=horizon, steps=T)
        with_returns = fuse_eval and not ppo.use_RND
        if with_returns:
            ops.gae_columns(buf.rewards, buf.dones, buf.values, buf.lengths, ppo.gamma, ppo.GAE_lambda, out=buf.returns)
        scores = self._scores.cpu().numpy()  # the one host sync of the rollout: reward sum, number of env steps
        n_new = int(scores[1])
        self.reward_score += scores[0]
        self.step_score += n_new
        ppo.memory.append_from_rollout(buf, n_new, scalar_actions=not d.is_continuous, eval_tag=ppo._eval_tag() if fuse_eval else None,
                                       with_returns=with_returns)

    def _worker_stepwise(self, initial_states=None):
        """The reference's loop, one kernel set per time step, host numpy between the calls.  Draws the same random
        numbers as the fused kernel (Philox keyed by env index, time step and episode), so both give the same buffer."""
        env, ppo = self.env, self.ppo
        states = env.reset()[0] if initial_states is None else env.reset_to_device(initial_states).cpu().numpy()
        ours = hasattr(ppo, "get_action_device")
        while True:
            if ours:
                idx, n = env.active_indices()
                actions = ppo.get_action(t.from_numpy(states), _row_ids=idx[:n].contiguous(), _call_index=(env.episode << 32) | env.t)
            else:
                actions = ppo.get_action(t.from_numpy(states))
            next_states, rewards, dones, truncates, _ = env.step(actions)
            finished = dones | truncates
            mask = env.envs_active
            utils.buffer_append(self.buffer, states, actions, rewards, finished, mask, self.num_envs)
            self.reward_score += np.sum(rewards)
            self.step_score += np.sum(~mask)
            states = utils.inactive_states_dropout(next_states, finished)
            env.envs_active = utils.update_active_environments_list(mask, finished)
            if np.all(env.envs_active):
                utils.buffer_to_target_buffer_transfer(self.buffer, ppo.memory)
                break

    def run(self):
        """AsyncPPO.py:148-165."""
        pbar = tqdm(total=self.steps, unit='step', disable=not self.show_progress)
        done_steps = 0
        while done_steps < self.steps:
            self.step_score = 0
            self.reward_score = 0

            self.worker()

            mean_reward = self.reward_score / self.num_envs
            inc = min(self.steps - done_steps, int(self.step_score))
            done_steps += inc
            pbar.update(inc)
            pbar.set_description(f'Mean reward {mean_reward: .1f}')

            self.ppo.learn()
        pbar.close()
        self.env.close()
