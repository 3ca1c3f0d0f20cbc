"""Generate the golden fixtures in tests/golden/*.npz by running the REAL reference code.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box).
    python tests/golden/gen_golden.py

What is executed verbatim from /root/reference (never copied into this repo):
  * AsyncTools/AsyncPPO.py  EnvVectorizer, VecMemory, AsyncPPO.worker   (:11-146)
  * AsyncTools/utils.py     all seven free functions                      (:1-50)
  * PPO/PPO.py              PPO.__init__, compute_gae, learn, get_action  (:13-260)
  * PPO/ActorCritic.py, PPO/RND.py, PPO/Memory.py

gymnasium is absent, so `import gymnasium` in AsyncTools/AsyncPPO.py:3 is satisfied by a stub module
exposing a bare `Env` class, and the per-env objects are oracle/envs.py (our restatement of the
classic-control physics; parity with gymnasium itself is unpinned).  Sampling is teacher-forced: a
fake `ppo` object replays a taped action per (t, env).
"""
from __future__ import annotations

import os
import sys
import types

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
os.environ["CUDA_VISIBLE_DEVICES"] = ""

import numpy as np
import torch as t

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import envs as oenvs  # noqa: E402


def import_reference():
    gym = types.ModuleType("gymnasium")

    class Env:  # the only attribute AsyncTools/AsyncPPO.py touches at import time
        pass

    gym.Env = Env
    sys.modules["gymnasium"] = gym
    sys.path.insert(0, REF)
    import AsyncTools  # noqa: F401
    import AsyncTools.AsyncPPO as ref_async
    import AsyncTools.utils as ref_utils
    import PPO as ref_ppo

    assert ref_async.__file__.startswith(REF) and ref_ppo.__file__.startswith(REF)
    return ref_async, ref_utils, ref_ppo


class TapedPolicy:
    """Duck-typed `ppo` for AsyncPPO (needs get_action / memory / learn): replays tape[t][env]."""

    def __init__(self, vec_env, tape, memory):
        self.vec_env, self.tape, self.memory, self.t = vec_env, tape, memory, 0
        self.seen_states = []

    def get_action(self, states):
        active = ~self.vec_env.envs_active
        self.seen_states.append(states.numpy().copy())
        a = self.tape[self.t][active]
        self.t += 1
        return a

    def learn(self):
        pass


def initial_states(env_id, E, rng):
    if env_id == "CartPole-v1":
        return rng.uniform(-0.05, 0.05, size=(E, 4))
    if env_id == "Pendulum-v1":
        hi = np.array([np.pi, 1.0])
        return rng.uniform(-hi, hi, size=(E, 2))
    if env_id == "MountainCarContinuous-v0":   # (float64 start states, as reset() leaves them: the first step computes in double)
        return rng.uniform([-1.2, -0.07], [0.5, 0.07], size=(E, 2))
    if env_id == "MountainCar-v0":   # wider than gymnasium's reset (-0.6..-0.4, 0) so that goals and the left wall are reached
        return rng.uniform([-1.2, -0.07], [0.55, 0.07], size=(E, 2))
    return rng.uniform(-0.1, 0.1, size=(E, 4)).astype(np.float32).astype(np.float64)


def make_tape(env_id, T, E, rng):
    if env_id == "Pendulum-v1":
        return (2.0 * np.tanh(rng.standard_normal((T, E, 1)))).astype(np.float32)
    if env_id == "MountainCarContinuous-v0":   # |a| up to 1.6: the force clamp (python-float branch) is exercised
        return (1.6 * np.tanh(rng.standard_normal((T, E, 1)))).astype(np.float32)
    n = 2 if env_id == "CartPole-v1" else 3
    return rng.integers(0, n, size=(T, E)).astype(np.int64)


def gen_rollout(ref_async, env_id, E, T, seed):
    rng = np.random.default_rng(seed)
    env = oenvs.make(env_id, max_episode_steps=T)
    s0 = initial_states(env_id, E, rng)
    tape = make_tape(env_id, T, E, rng)
    runner = ref_async.AsyncPPO(env=env, ppo=None, num_envs=E, steps=10)
    for i in range(E):
        runner.env.envs[i].inject_state(s0[i])
    from PPO import Memory

    pol = TapedPolicy(runner.env, tape, Memory())
    runner.ppo = pol
    # trace every vectorised step without touching the reference code
    trace = dict(obs=[], rewards=[], dones=[], truncs=[], mask_before=[])
    orig_step = runner.env.step

    def traced_step(actions):
        trace["mask_before"].append(runner.env.envs_active.copy())
        out = orig_step(actions)
        trace["obs"].append(out[0].copy()); trace["rewards"].append(np.asarray(out[1], np.float64).copy())
        trace["dones"].append(out[2].copy()); trace["truncs"].append(out[3].copy())
        return out

    runner.env.step = traced_step
    runner.step_score = 0
    runner.reward_score = 0
    runner.worker()
    mem = pol.memory
    nsteps = len(trace["obs"])
    out = dict(
        env_id=np.array(env_id), E=E, T=T, init_state=s0, tape=tape[:nsteps],
        states=np.array(mem.states, np.float32), actions=np.array(mem.actions, np.float32),
        rewards=np.array(mem.rewards, np.float32), dones=np.array(mem.dones, np.float32),
        reward_score=np.float64(runner.reward_score), step_score=np.int64(runner.step_score),
        nsteps=nsteps,
        mask_before=np.stack(trace["mask_before"]),
        final_state=np.stack([np.asarray(e.state, np.float64) for e in runner.env.envs]),
    )
    # ragged per-step compact arrays, stored concatenated with offsets
    out["step_counts"] = np.array([len(r) for r in trace["rewards"]], np.int64)
    out["step_obs"] = np.concatenate(trace["obs"], 0)
    out["step_rewards"] = np.concatenate(trace["rewards"], 0)
    out["step_dones"] = np.concatenate(trace["dones"], 0)
    out["step_truncs"] = np.concatenate(trace["truncs"], 0)
    out["seen_states"] = np.concatenate(pol.seen_states, 0)
    return out


def flat_params(module):
    return np.concatenate([p.detach().cpu().numpy().ravel() for p in module.parameters()]).astype(np.float32)


def sd_arrays(prefix, module):
    return {f"{prefix}{k}": v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def gen_learn(ref_ppo, roll, *, is_continuous, O, A, seed, use_rnd=False, action_scaling=None,
              k_epochs=3, mini_batch_size=48, batch_size=64, lr=1e-3):
    t.manual_seed(seed)
    ppo = ref_ppo.PPO(is_continuous=is_continuous, observ_dim=O, action_dim=A, action_scaling=action_scaling,
                      lr=lr, k_epochs=k_epochs, batch_size=batch_size, mini_batch_size=mini_batch_size,
                      use_RND=use_rnd, beta=0.001)
    out = dict(is_continuous=is_continuous, O=O, A=A, k_epochs=k_epochs, mini_batch_size=mini_batch_size,
               batch_size=batch_size, lr=lr, gamma=ppo.gamma, GAE_lambda=ppo.GAE_lambda,
               policy_clip=ppo.policy_clip, use_rnd=use_rnd, beta=0.001)
    out.update(sd_arrays("init.", ppo.policy))
    out["init_flat"] = flat_params(ppo.policy)
    if use_rnd:
        out.update(sd_arrays("rnd_init.", ppo.rnd))
    states, actions = roll["states"], roll["actions"]
    # forward-only facts about the initial policy
    with t.no_grad():
        st = t.from_numpy(states)
        ac = t.from_numpy(actions)
        logp, val, ent = ppo.policy_old.get_evaluate(st, ac)
        out["eval_logp"], out["eval_value"], out["eval_entropy"] = logp.numpy(), val.numpy(), ent.numpy()
        dist = ppo.policy_old.get_dist(st)
        if is_continuous:
            out["dist_mu"] = dist.loc.numpy()
            out["dist_std"] = dist.scale_tril.diagonal(dim1=-2, dim2=-1).numpy()
        else:
            out["dist_probs"] = dist.probs.numpy()
        if use_rnd:
            out["rnd_intrinsic"] = ppo.rnd.compute_intrinsic_reward(ppo.batch_packer(st, mini_batch_size)).numpy()
    # fill PPO.memory exactly as buffer_to_target_buffer_transfer would (lists of float32 items)
    for i in range(len(states)):
        ppo.memory.states.append(states[i].copy())
        ppo.memory.actions.append(actions[i].copy())
        ppo.memory.rewards.append(np.float32(roll["rewards"][i]))
        ppo.memory.dones.append(np.float32(roll["dones"][i]))
    rec = {}
    orig_gae = ppo.compute_gae

    def rec_gae(rewards, dones, state_values, next_value):
        ret = orig_gae(rewards, dones, state_values, next_value)
        rec.update(gae_rewards=np.array(rewards, np.float32), gae_dones=np.array(dones, np.float32),
                   gae_values=np.array(state_values, np.float32), gae_next_value=np.float32(next_value),
                   gae_returns=np.array(ret, np.float32))
        return ret

    ppo.compute_gae = rec_gae
    # per-minibatch loss means, read back from the reference's own tqdm label (PPO.py:254-255)
    ppo.learn()
    out.update(rec)
    ret = t.from_numpy(rec["gae_returns"]); v = t.from_numpy(rec["gae_values"])
    adv = t.sub(ret, v)
    out["advantages"] = ((adv - adv.mean()) / (adv.std() + 1e-8)).numpy()
    out.update(sd_arrays("post.", ppo.policy))
    out["post_flat"] = flat_params(ppo.policy)
    opt_state = ppo.optimizer.state_dict()["state"]
    out["post_exp_avg"] = np.concatenate([opt_state[i]["exp_avg"].numpy().ravel() for i in sorted(opt_state)])
    out["post_exp_avg_sq"] = np.concatenate([opt_state[i]["exp_avg_sq"].numpy().ravel() for i in sorted(opt_state)])
    if use_rnd:
        out.update(sd_arrays("rnd_post.", ppo.rnd))
    assert len(ppo.memory.states) == 0
    return out


def gen_single_step(ref_ppo, roll, *, is_continuous, O, A, seed):
    """One optimiser step on one minibatch = the whole buffer: pins loss and gradient numerics."""
    t.manual_seed(seed)
    N = len(roll["states"])
    return gen_learn(ref_ppo, roll, is_continuous=is_continuous, O=O, A=A, seed=seed, k_epochs=1,
                     mini_batch_size=N, batch_size=8, action_scaling=2.0 if is_continuous else None)


def gen_utils(ref_async, ref_utils, seed):
    rng = np.random.default_rng(seed)
    out = {}
    E = 11
    mask = rng.random(E) < 0.4
    mask[3] = False
    n = int((~mask).sum())
    dones = rng.random(n) < 0.5
    states = rng.standard_normal((n, 4)).astype(np.float32)
    out["mask"] = mask.copy(); out["dones"] = dones.copy(); out["states"] = states
    out["indexes"] = ref_utils.indexes_of_active_environments(E, mask)
    out["number"] = np.int64(ref_utils.number_of_active_environments(mask))
    out["range"] = ref_utils.range_of_active_environments(mask)
    out["dropout"] = ref_utils.inactive_states_dropout(states, dones)
    m2 = mask.copy()
    r = ref_utils.update_active_environments_list(m2, dones)
    assert r is m2
    out["mask_after"] = m2
    # buffer_append x3 then transfer
    buf = ref_async.VecMemory(E)
    from PPO import Memory
    cur = np.zeros(E, bool)
    steps = []
    for _ in range(4):
        n = int((~cur).sum())
        if n == 0:
            break
        s = rng.standard_normal((n, 3)); a = rng.integers(0, 3, n); rw = rng.standard_normal(n); d = rng.random(n) < 0.45
        ref_utils.buffer_append(buf, s, a, rw, d, cur, E)
        steps.append((s, a, rw, d, cur.copy()))
        ref_utils.update_active_environments_list(cur, d)
    mem = Memory()
    ref_utils.buffer_to_target_buffer_transfer(buf, mem)
    out["ba_nsteps"] = len(steps)
    for i, (s, a, rw, d, m) in enumerate(steps):
        out[f"ba_s{i}"], out[f"ba_a{i}"], out[f"ba_r{i}"], out[f"ba_d{i}"], out[f"ba_m{i}"] = s, a, rw, d, m
    out["ba_states"] = np.array(mem.states, np.float32); out["ba_actions"] = np.array(mem.actions, np.float32)
    out["ba_rewards"] = np.array(mem.rewards, np.float32); out["ba_dones"] = np.array(mem.dones, np.float32)
    assert all(len(x) == 0 for x in buf.states)
    return out


def main():
    ref_async, ref_utils, ref_ppo = import_reference()
    t.set_num_threads(1)
    if "--only-mountaincar" in sys.argv:   # added after the other fixtures were frozen
        v = gen_rollout(ref_async, "MountainCar-v0", E=16, T=60, seed=14)
        np.savez_compressed(os.path.join(HERE, "rollout_mountaincar.npz"), **v)
        print("mountaincar N =", len(v["states"]), "steps", v["nsteps"], "reward", float(v["reward_score"]), "goals", int((v["dones"] > 0).sum()))
        return
    if "--only-mountaincarcont" in sys.argv:   # added in round 2
        v = gen_rollout(ref_async, "MountainCarContinuous-v0", E=24, T=120, seed=15)
        np.savez_compressed(os.path.join(HERE, "rollout_mountaincarcont.npz"), **v)
        print("mountaincarcont N =", len(v["states"]), "steps", v["nsteps"], "reward", float(v["reward_score"]), "episode ends", int((v["dones"] > 0).sum()))
        return
    rolls = {
        "cartpole": gen_rollout(ref_async, "CartPole-v1", E=12, T=40, seed=11),
        "pendulum": gen_rollout(ref_async, "Pendulum-v1", E=6, T=25, seed=12),
        "acrobot": gen_rollout(ref_async, "Acrobot-v1", E=6, T=30, seed=13),
    }
    for k, v in rolls.items():
        np.savez_compressed(os.path.join(HERE, f"rollout_{k}.npz"), **v)
        print(k, "N =", len(v["states"]), "steps", v["nsteps"], "reward", float(v["reward_score"]))
    learn = {
        "discrete": gen_learn(ref_ppo, rolls["cartpole"], is_continuous=False, O=4, A=2, seed=0),
        "continuous": gen_learn(ref_ppo, rolls["pendulum"], is_continuous=True, O=3, A=1, seed=1, action_scaling=2.0),
        "rnd": gen_learn(ref_ppo, rolls["acrobot"], is_continuous=False, O=6, A=3, seed=2, use_rnd=True),
        "discrete_1step": gen_single_step(ref_ppo, rolls["cartpole"], is_continuous=False, O=4, A=2, seed=3),
        "continuous_1step": gen_single_step(ref_ppo, rolls["pendulum"], is_continuous=True, O=3, A=1, seed=4),
    }
    for k, v in learn.items():
        np.savez_compressed(os.path.join(HERE, f"learn_{k}.npz"), **v)
        print("learn", k, "params", v["post_flat"].size, "max|dW|", float(np.abs(v["post_flat"] - v["init_flat"]).max()))
    np.savez_compressed(os.path.join(HERE, "utils.npz"), **gen_utils(ref_async, ref_utils, 21))
    print("done")


if __name__ == "__main__":
    main()
