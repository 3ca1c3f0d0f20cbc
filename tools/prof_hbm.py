"""Profiling driver (GPU box): one pass of each HBM-bound kernel of the path at the C2 shapes (E = 65 536, T = 128):
flat GAE, column GAE, advantage normalisation, buffer transfer, standalone env step, taped rollout."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-reinforcement-learning_b200")]
import torch as t
from prl_b200 import ops

E, T = 65536, 128
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = t.device("cuda", 0)
N = E * T
r = t.ones(N, device=dev); v = t.rand(N, device=dev); ret = t.empty(N, device=dev); adv = t.empty(N, device=dev)
d = t.zeros(N, device=dev); d[T - 1::T] = 1.0
lens_rag = t.randint(20, T + 1, (2 * E,), device=dev)          # ragged episodes (20..128 steps) covering all N transitions
ends_rag = t.cumsum(lens_rag, 0) - 1
drag = t.zeros(N, device=dev); drag[ends_rag[ends_rag < N]] = 1.0; drag[N - 1] = 1.0
d2 = t.zeros(T, E, device=dev); d2[T - 1] = 1.0
lens = t.full((E,), T, dtype=t.int32, device=dev)
stats = t.zeros(4, dtype=t.float64, device=dev)
sim = ops.EnvState("CartPole-v1", E, T)
buf = ops.RolloutBuffer(E, T, 4, 1)
tape = t.randint(0, 2, (T, E), dtype=t.int32, device=dev)
scores = t.zeros(2, dtype=t.float64, device=dev)
mem = [t.empty(N, 4, device=dev), t.empty(N, 1, device=dev), t.empty(N, device=dev), t.empty(N, device=dev)]
total = t.zeros(1, dtype=t.int64, device=dev)
E2 = 1 << 20
sim2 = ops.EnvState("CartPole-v1", E2, 1 << 30); sim2.reset(1, 1)
idx = t.arange(E2, dtype=t.int32, device=dev); acts = t.randint(0, 2, (E2,), dtype=t.int32, device=dev)
flush = t.empty(256 << 20, dtype=t.uint8, device=dev)


def timed(name, fn, nbytes):
    ms = 0.0
    fn(); t.cuda.synchronize()   # warm-up (module load, attribute calls)
    for _ in range(reps):
        flush.fill_(0)
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); t.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    ms /= reps
    print(f"{name:28s} {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:8.1f} GB/s")


def taped():
    sim.reset(1, 1); scores.zero_()
    ops.rollout(sim, buf, None, 1.0, 0, 1, scores, tape=tape)


def filled_transfer():
    buf.lengths.fill_(T)
    buf.transfer(*mem, 0, total)


ws = t.empty(64, dtype=t.uint8, device=dev)
timed("gae_flat_T128", lambda: ops.gae(r, d, v, 0.995, 0.95, out=ret, ws=ws), N * 16)
timed("gae_flat_ragged", lambda: ops.gae(r, drag, v, 0.995, 0.95, out=ret, ws=ws), N * 16)
timed("gae_columns_T128", lambda: ops.gae_columns(r.view(T, E), d2, v.view(T, E), lens, 0.995, 0.95, out=ret.view(T, E)), N * 16)
timed("adv_normalize", lambda: (stats.zero_(), ops.adv_normalize(ret, v, stats=stats, phase=3, out=adv)), N * 20)
timed("buffer_transfer_full", filled_transfer, N * 56)
timed("env_step_cartpole_1M", lambda: sim2.step(idx, E2, acts), E2 * 102)
timed("rollout_taped", taped, 0)
