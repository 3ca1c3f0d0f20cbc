"""Scratch (GPU box): wall-clock breakdown of one bench step, device-reset vs host-injected start states."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-reinforcement-learning_b200")]
import numpy as np, torch as t
import prl_b200
from AsyncTools.AsyncPPO import AsyncPPO
from PPO import PPO

E, T = 65536, 128
t.manual_seed(0)
ppo = PPO(is_continuous=False, observ_dim=4, action_dim=2, lr=1e-3, k_epochs=11, batch_size=1024, mini_batch_size=65536)
ppo.show_progress = False
ppo.use_cuda_graph = True
ap = AsyncPPO(env=prl_b200.make("CartPole-v1", max_episode_steps=T), ppo=ppo, num_envs=E, steps=1)
host_states = t.from_numpy(np.random.default_rng(0).uniform(-0.05, 0.05, (E, 4))).pin_memory()
def sync(): t.cuda.synchronize(); return time.perf_counter()
for mode in ["dev"] * 10 + ["host"] * 6 + ["dev"] * 4:
    t0 = sync()
    ap.step_score = 0; ap.reward_score = 0
    ap.worker(initial_states=host_states if mode == "host" else None)
    t1 = sync()
    ppo.learn()
    t2 = time.perf_counter()
    t3 = sync()
    print(f"{mode}: worker {1e3*(t1-t0):7.1f} ms  learn host {1e3*(t2-t1):7.1f} ms  learn gpu-tail {1e3*(t3-t2):7.1f} ms  N={int(ap.step_score)}")
