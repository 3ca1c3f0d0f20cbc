"""Profiling driver (GPU box): a few launches of the PPO minibatch-gradient kernels on 65 536 CartPole-shaped rows."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-reinforcement-learning_b200")]
import numpy as np, torch as t
from prl_b200 import ops

path = sys.argv[1] if len(sys.argv) > 1 else "tc"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
O, A = (int(x) for x in os.environ.get("PRL_PROF_SHAPE", "4,2").split(","))   # observ_dim, action_dim (discrete policy)
rng = np.random.default_rng(0)
if (O, A) == (4, 2):
    params = t.from_numpy(np.load(os.path.join(ROOT, "tests/golden/learn_discrete.npz"))["init_flat"]).cuda()
else:   # a freshly initialised reference-style network of that shape
    sys.path.insert(0, os.path.join(ROOT, "parallel-reinforcement-learning_b200"))
    from PPO import ActorCritic
    t.manual_seed(0)
    params = ActorCritic(False, O, A).flat.detach().clone()
s = t.from_numpy(rng.uniform(-1, 1, (N, O)).astype(np.float32)).cuda()
a = t.from_numpy(rng.integers(0, A, (N, 1)).astype(np.float32)).cuda()
logp, _, _ = ops.policy_evaluate(params, False, O, A, s, a)
adv = t.randn(N, device="cuda"); ret = t.randn(N, device="cuda")
grad = t.zeros_like(params); loss = t.zeros(4, dtype=t.float64, device="cuda")
if path == "step":   # fused gradient + clip + AdamW (cooperative launch), eager and as a replayed CUDA graph
    from prl_b200.optim import FusedAdamW
    opt = FusedAdamW(params, 1e-4)
    ws = t.zeros(ops.update_tc_ws_floats(False, O, A, N), device="cuda")
    fn = lambda: ops.ppo_step_tc(params, False, O, A, s, a, logp, adv, ret, 0.2, 1.0 / N, grad, loss, opt, ws)
    if os.environ.get("PRL_PROF_STREAM"):
        # every launch of the replayed graph works on its OWN minibatch: 96 x N rows (200 MB at N = 65 536, more than the L2), as
        # inside PPO.learn - the launches of the plain graph mode below re-read one L2-resident minibatch
        K = 96
        S = t.from_numpy(rng.uniform(-1, 1, (K * N, O)).astype(np.float32)).cuda()
        Aa = t.from_numpy(rng.integers(0, A, (K * N, 1)).astype(np.float32)).cuda()
        LP, _, _ = ops.policy_evaluate(params, False, O, A, S, Aa)
        AD = t.randn(K * N, device="cuda"); RT = t.randn(K * N, device="cuda")
        side = t.cuda.Stream(); side.wait_stream(t.cuda.current_stream())
        gr = t.cuda.CUDAGraph()
        fn()
        with t.cuda.stream(side):
            gr.capture_begin()
            for k in range(K):
                sl = slice(k * N, (k + 1) * N)
                ops.ppo_step_tc(params, False, O, A, S[sl], Aa[sl], LP[sl], AD[sl], RT[sl], 0.2, 1.0 / N, grad, loss, opt, ws)
            gr.capture_end()
        t.cuda.current_stream().wait_stream(side)
        fn = gr.replay
        N = N * K
    elif os.environ.get("PRL_PROF_GRAPH"):
        side = t.cuda.Stream(); side.wait_stream(t.cuda.current_stream())
        gr = t.cuda.CUDAGraph()
        with t.cuda.stream(side):
            gr.capture_begin()
            for _ in range(32):
                fn()
            gr.capture_end()
        t.cuda.current_stream().wait_stream(side)
        one = fn
        fn = gr.replay
        N = N * 32
elif path == "tc":
    ws = t.zeros(ops.update_tc_ws_floats(False, O, A, N), device="cuda")
    fn = lambda: ops.ppo_grad_tc(params, False, O, A, s, a, logp, adv, ret, 0.2, 1.0 / N, grad, loss, ws)
else:
    ws = t.zeros(ops.update_ws_floats(False, O, A, N), device="cuda")
    fn = lambda: ops.ppo_grad(params, False, O, A, s, a, logp, adv, ret, 0.2, 1.0 / N, grad, loss, ws)
for _ in range(3):
    fn()
t.cuda.synchronize()
e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    fn()
e1.record(); t.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{path} N={N}: {ms * 1e3:.1f} us per call, {N / ms / 1e3:.1f} M rows/s, {N * 51840 / ms / 1e9:.2f} TFLOP/s algorithmic")
if os.environ.get("PRL_PROF_EVAL"):
    NE = 8_388_608
    se = t.from_numpy(rng.uniform(-1, 1, (NE, O)).astype(np.float32)).cuda()
    ae = t.from_numpy(rng.integers(0, A, (NE, 1)).astype(np.float32)).cuda()
    lo, vo = t.empty(NE, device="cuda"), t.empty(NE, device="cuda")
    ent = t.zeros(1, dtype=t.float64, device="cuda")
    for _ in range(2):
        ops.policy_evaluate(params, False, O, A, se, ae, entropy_sum=ent, logp=lo, value=vo)
    t.cuda.synchronize()
    e0.record()
    for _ in range(5):
        ops.policy_evaluate(params, False, O, A, se, ae, entropy_sum=ent, logp=lo, value=vo)
    e1.record(); t.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"evaluate N={NE}: {ms:.3f} ms, {NE / ms / 1e3:.1f} M rows/s, {NE * 17280 / ms / 1e9:.2f} TFLOP/s fp32")
