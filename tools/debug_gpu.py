"""Scratch diagnostics run on the GPU box (not a test): per-block gradient error vs an fp64 oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-reinforcement-learning_b200")]
import numpy as np, torch as t
from oracle import ppo as oppo
from prl_b200 import ops

G = os.path.join(ROOT, "tests", "golden")
dev = lambda a: t.from_numpy(np.ascontiguousarray(a)).cuda()
for name, roll in [("discrete_1step", "cartpole"), ("continuous_1step", "pendulum")]:
    g = np.load(f"{G}/learn_{name}.npz"); r = np.load(f"{G}/rollout_{roll}.npz")
    cont, O, A = bool(g["is_continuous"]), int(g["O"]), int(g["A"])
    N = len(r["states"])
    params = dev(g["init_flat"]).clone()
    s = dev(r["states"]); a = dev(r["actions"].reshape(N, -1))
    logp, val, _ = ops.policy_evaluate(params, cont, O, A, s, a)
    print(name, "old logp err", np.abs(logp.cpu().numpy() - g["eval_logp"]).max())
    for which, old in (("device old_logp", logp), ("reference old_logp", dev(g["eval_logp"]))):
        grad = t.zeros_like(params); loss = t.zeros(4, dtype=t.float64, device="cuda")
        ws = t.empty(ops.update_ws_floats(cont, O, A, N), device="cuda")
        ops.ppo_grad(params, cont, O, A, s, a, old, dev(g["advantages"]), dev(g["gae_returns"]), float(g["policy_clip"]), 1.0 / N, grad, loss, ws)
        got = grad.cpu().numpy().astype(np.float64)
        keys = oppo.param_keys(cont)
        def ograd(dtype):
            p = {k: v.to(dtype).requires_grad_(True) for k, v in oppo.unflatten(g["init_flat"], cont, O, A).items()}
            c = lambda x: t.from_numpy(np.asarray(x)).to(dtype)
            lo = oppo.ppo_loss(p, cont, c(r["states"]), c(r["actions"]), c(g["eval_logp"]), c(g["advantages"]), c(g["gae_returns"]), float(g["policy_clip"]))
            return [x.reshape(-1).double().numpy() for x in t.autograd.grad(lo, [p[k] for k in keys])]
        w64, w32 = ograd(t.float64), ograd(t.float32)
        off = 0
        print(" ", which)
        for k, a64, a32 in zip(keys, w64, w32):
            n = len(a64); mine = got[off:off + n]; off += n
            sc = np.abs(a64).max() + 1e-30
            print(f"    {k:24s} max|g|={sc:.3e} mine-vs-f64={np.abs(mine - a64).max() / sc:.2e} torch32-vs-f64={np.abs(a32 - a64).max() / sc:.2e}")
