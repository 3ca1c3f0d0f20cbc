"""Two-rank GPU worker (launched by tests/test_gpu_multi.py through torch.distributed.run): the sharded PPO update with
the in-kernel peer-memory gradient exchange (prl_ppo_step_tc_p2p) must leave every rank with the same weights as the
NCCL-allreduce path, eager and as a CUDA graph.  Writes `ok` to <out>/rank<r> on success."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "parallel-reinforcement-learning_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
import torch as t

from prl_b200 import dist, make
from PPO import PPO
from AsyncTools.AsyncPPO import AsyncPPO

out = sys.argv[1]
comm = dist.init_from_env()
rank = comm.rank
env = make("CartPole-v1")


def gather2(x):
    """[rank 0's x, rank 1's x] on the host (through the CPU, so that it also works over gloo when both ranks share one GPU)."""
    both = [None, None]
    t.distributed.all_gather_object(both, x.detach().cpu())
    return both



def run(peer_exchange, graph, envs):
    t.manual_seed(1234)
    ppo = PPO(is_continuous=False, observ_dim=4, action_dim=2, lr=3e-4, k_epochs=3, mini_batch_size=1024, batch_size=256, use_RND=False)
    ppo.update_path, ppo.peer_exchange, ppo.use_cuda_graph, ppo.graph_collectives, ppo.show_progress = "tensor", peer_exchange, graph, graph, False
    t.manual_seed(99 + rank)   # a different env reset stream per shard
    a = AsyncPPO(env=env, ppo=ppo, num_envs=envs, steps=10**9)
    for _ in range(2):   # two rollouts + updates: the second one exercises the persistent exchange flags / step parity
        a.worker()
        ppo.learn()
    t.cuda.synchronize()
    return ppo.policy.flat.clone(), None


# rank 1 gets fewer envs, so its row count differs and the last minibatches have no rows on it (b == 0 launches)
envs = 192 if rank == 0 else 40
ref, _ = run(False, False, envs)
for graph in (False, True):
    got, st = run(True, graph, envs)
    d = (got - ref).abs().max().item()
    # rank-order summation of two addends is commutative, so the exchange is bit-identical to NCCL's sum
    assert d == 0.0, f"rank {rank} graph={graph}: peer-exchange weights differ from NCCL path by {d}"
    both = gather2(got)
    assert t.equal(both[0], both[1]), "ranks diverged"
# Acrobot + RND, sharded: the predictor update allreduces its re-weighted gradient, so the replicas stay identical
t.manual_seed(7)
ppo = PPO(is_continuous=False, observ_dim=6, action_dim=3, lr=3e-4, k_epochs=2, mini_batch_size=2048, batch_size=256, use_RND=True, beta=0.001)
ppo.show_progress = False
pred0 = ppo.rnd.pred_flat.clone()
t.manual_seed(5 + rank)
a = AsyncPPO(env=make("Acrobot-v1", max_episode_steps=64), ppo=ppo, num_envs=48 if rank == 0 else 16, steps=10**9)
a.worker()
ppo.learn()
t.cuda.synchronize()
for flat in (ppo.rnd.pred_flat, ppo.policy.flat):
    both = gather2(flat)
    assert t.equal(both[0], both[1]), "RND run: ranks diverged"
assert not t.equal(pred0, ppo.rnd.pred_flat) and t.isfinite(ppo.rnd.pred_flat).all()

# ---- the sharded update against the ORACLE fed the same permutation.  The ranks are seeded DIFFERENTLY on purpose:
# PPO.sync_replicas (rank 0's networks / optimiser state broadcast at construction) must make them start identical, and
# rank_seed must give the shards different action noise and reset streams.
from oracle import ppo as oppo  # noqa: E402  (the checker)

t.manual_seed(100 + 17 * rank)
ppo = PPO(is_continuous=False, observ_dim=4, action_dim=2, lr=1e-3, k_epochs=2, mini_batch_size=512, batch_size=256)
ppo.show_progress = False
a = AsyncPPO(env=make("CartPole-v1", max_episode_steps=48), ppo=ppo, num_envs=64 if rank == 0 else 40, steps=10**9)
a.worker()
ms, ma, mr, md = (x.cpu().numpy() for x in ppo.memory.device_view(4, 1, ppo.device))
init = ppo.policy.flat.cpu().numpy().copy()
shards = [None, None]
t.distributed.all_gather_object(shards, dict(states=ms, actions=ma[:, 0], rewards=mr, dones=md, init=init, seed=ppo._seed))
assert np.array_equal(shards[0]["init"], shards[1]["init"]), "sync_replicas left the replicas different"
assert shards[0]["seed"] != shards[1]["seed"]
assert not np.array_equal(shards[0]["states"][:16], shards[1]["states"][:16]), "both shards reset identically"
ppo.learn()
t.cuda.synchronize()
got = ppo.policy.flat.cpu().numpy()
n_all = [len(sh["rewards"]) for sh in shards]
mb_local, n_mb, counts = dist.minibatch_schedule(n_all, 512)
offs = [0, n_all[0]]
schedule = [np.concatenate([offs[r] + np.arange(min(k * mb_local, n_all[r]), min((k + 1) * mb_local, n_all[r])) for r in range(2)]) for k in range(n_mb)]
assert [len(x) for x in schedule] == counts and sorted(np.concatenate(schedule).tolist()) == list(range(sum(n_all)))
mem = {k: np.concatenate([sh[k] for sh in shards]) for k in ("states", "actions", "rewards", "dones")}
p32 = oppo.unflatten(init, False, 4, 2)
oppo.learn(p32, False, mem, lr=1e-3, k_epochs=2, policy_clip=0.2, gae_lambda=0.95, gamma=0.995, mini_batch_size=512, schedule=schedule)
want = oppo.flatten(p32, False).numpy()
d = np.abs(got.astype(np.float64) - want)
ok = d <= 1e-5 * np.abs(want) + 2e-6
print(f"[parity] rank {rank}: sharded learn() ({sum(n_all)} rows on 2 ranks as {n_all}, {2 * n_mb} optimiser steps, in-kernel peer exchange) vs oracle.ppo.learn "
      f"fed the union-of-k-th-chunks schedule: {100 * ok.mean():.2f} % within 1e-5*|w| + 2e-6, max |diff| {d.max():.2e}", flush=True)
assert ok.mean() >= 0.9995 and d.max() <= 1e-5, (ok.mean(), d.max())   # achieved 99.99 %, 4.1e-6
both = gather2(ppo.policy.flat)
assert t.equal(both[0], both[1]), "H7 run: ranks diverged"
ppo.close()
open(os.path.join(out, f"rank{rank}"), "w").write("ok")
t.distributed.destroy_process_group()
