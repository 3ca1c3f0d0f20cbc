"""Profiling driver (GPU box, under torch.distributed.run with >= 2 ranks): the sharded PPO optimiser step with the
in-kernel peer-memory gradient exchange (prl_ppo_step_tc_p2p) on 65 536 CartPole-shaped rows per rank, eager (with
PRL_TC_TIMING phase stamps for a few launches) and as a replayed CUDA graph of 32 launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-reinforcement-learning_b200")]
import numpy as np, torch as t
from prl_b200 import dist, ops
from prl_b200.optim import FusedAdamW

comm = dist.init_from_env()
rank, world = comm.rank, comm.world_size
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
O, A = 4, 2
g = np.load(os.path.join(ROOT, "tests/golden/learn_discrete.npz"))
rng = np.random.default_rng(rank)
params = t.from_numpy(g["init_flat"]).cuda()
s = t.from_numpy(rng.uniform(-1, 1, (N, O)).astype(np.float32)).cuda()
a = t.from_numpy(rng.integers(0, A, (N, 1)).astype(np.float32)).cuda()
logp, _, _ = ops.policy_evaluate(params, False, O, A, s, a)
adv = t.randn(N, device="cuda"); ret = t.randn(N, device="cuda")
grad = t.zeros_like(params); loss = t.zeros(4, dtype=t.float64, device="cuda")
opt = FusedAdamW(params, 1e-4)
ws = t.zeros(ops.update_tc_ws_floats(False, O, A, N), device="cuda")
xch = dist.PeerExchange(comm, False, O, A)
fn = lambda: ops.ppo_step_tc_p2p(params, False, O, A, s, a, logp, adv, ret, 0.2, 1.0 / (N * world), grad, loss, opt, ws, xch)
for _ in range(3):
    fn()
t.cuda.synchronize(); comm.barrier()
side = t.cuda.Stream(); side.wait_stream(t.cuda.current_stream())
gr = t.cuda.CUDAGraph()
os.environ.pop("PRL_TC_TIMING", None)
with t.cuda.stream(side):
    gr.capture_begin()
    for _ in range(32):
        fn()
    gr.capture_end()
t.cuda.current_stream().wait_stream(side)
gr.replay(); t.cuda.synchronize(); comm.barrier()
e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    gr.replay()
e1.record(); t.cuda.synchronize()
ms = e0.elapsed_time(e1) / (5 * 32)
print(f"rank {rank}/{world}: p2p step N={N}/rank: {ms * 1e3:.1f} us per launch in graph replay, {N * world / ms / 1e3:.1f} M rows/s aggregate", flush=True)
comm.barrier()
t.distributed.destroy_process_group()
