"""GPU box: one rollout + learn() at the full sizes of BASELINE configs[2] (Pendulum-v1, 262 144 envs, T = 200, tanh-Gaussian
policy) and configs[3] (Acrobot-v1 + RND, 65 536 envs), through the drop-in API.  Prints timings and basic invariants."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-reinforcement-learning_b200")]
import torch as t
import prl_b200
from PPO import PPO
from AsyncTools.AsyncPPO import AsyncPPO


def run(name, env_id, E, T, **kw):
    t.manual_seed(0)
    env = prl_b200.make(env_id, max_episode_steps=T)
    ppo = PPO(observ_dim=env.observ_dim, action_dim=env.action_dim, batch_size=1024, **kw)
    ppo.show_progress = False
    ap = AsyncPPO(env=env, ppo=ppo, num_envs=E, steps=1)
    before = ppo.policy.flat.clone()
    for it in range(2):
        t.cuda.synchronize(); t0 = time.time()
        ap.step_score = 0; ap.reward_score = 0
        ap.worker()
        t.cuda.synchronize(); t1 = time.time()
        n = len(ppo.memory.states)
        ppo.learn()
        t.cuda.synchronize(); t2 = time.time()
        print(f"{name} iter {it}: {n} transitions ({n / E:.1f} per env), rollout {1e3 * (t1 - t0):.1f} ms, learn {1e3 * (t2 - t1):.1f} ms, "
              f"{n / (t2 - t0) / 1e6:.2f} M env-steps/s, mean reward/step {float(ap.reward_score) / max(n, 1):.4f}", flush=True)
        assert n >= E and len(ppo.memory.states) == 0
    assert t.isfinite(ppo.policy.flat).all() and not t.equal(before, ppo.policy.flat)
    if kw.get("use_RND"):
        assert t.isfinite(ppo.rnd.pred_flat).all()
    print(f"{name}: ok, peak memory {t.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)


run("C3 Pendulum 262144 x 200", "Pendulum-v1", 262144, 200, is_continuous=True, action_scaling=2.0, k_epochs=3, mini_batch_size=262144)
run("C4 Acrobot+RND 65536 x 128", "Acrobot-v1", 65536, 128, is_continuous=False, use_RND=True, beta=0.001, k_epochs=3, mini_batch_size=65536)
run("MountainCar 65536 x 200", "MountainCar-v0", 65536, 200, is_continuous=False, k_epochs=3, mini_batch_size=65536)
