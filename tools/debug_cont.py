"""GPU box: per-parameter-block error of the continuous-policy gradient (tensor-core path vs fp32-FMA path vs torch float32) against
the float64 oracle autograd gradient, on the learn_continuous fixture rows with noisy old log-probs (the kernel test's setting)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-reinforcement-learning_b200")]
import numpy as np, torch as t
from prl_b200 import ops
from oracle import ppo as oppo
g = np.load(os.path.join(ROOT, "tests/golden/learn_continuous.npz")); r = np.load(os.path.join(ROOT, "tests/golden/rollout_pendulum.npz"))
cont, O, A = True, int(g["O"]), int(g["A"])
N = len(r["states"])
dev = lambda a: t.from_numpy(np.ascontiguousarray(a)).cuda()
params = dev(g["init_flat"]); s = dev(r["states"]); a = dev(r["actions"].reshape(N, -1))
rng = np.random.default_rng(11)
old_lp = g["eval_logp"] + rng.normal(0, 0.05, N).astype(np.float32)
adv_np = rng.standard_normal(N).astype(np.float32); ret_np = rng.standard_normal(N).astype(np.float32)
keys = oppo.param_keys(cont)
def oracle_grad(dtype):
    p = {k: v.to(dtype).requires_grad_(True) for k, v in oppo.unflatten(g["init_flat"], cont, O, A).items()}
    c = lambda x: t.from_numpy(np.asarray(x)).to(dtype)
    lo = oppo.ppo_loss(p, cont, c(r["states"]), c(r["actions"]), c(old_lp), c(adv_np), c(ret_np), 0.2)
    return lo.detach(), t.cat([x.reshape(-1) for x in t.autograd.grad(lo, [p[k] for k in keys])]).double().numpy()
lo, want = oracle_grad(t.float64); _, want32 = oracle_grad(t.float32)
grads = {"torch32": want32}
for path in ("tc", "fp32"):
    grad = t.full_like(params, float("nan")); loss = t.zeros(4, dtype=t.float64, device="cuda")
    if path == "tc":
        ws = t.zeros(ops.update_tc_ws_floats(cont, O, A, N), device="cuda")
        ops.ppo_grad_tc(params, cont, O, A, s, a, dev(old_lp), dev(adv_np), dev(ret_np), 0.2, 1.0 / N, grad, loss, ws)
    else:
        ws = t.empty(ops.update_ws_floats(cont, O, A, N), device="cuda")
        ops.ppo_grad(params, cont, O, A, s, a, dev(old_lp), dev(adv_np), dev(ret_np), 0.2, 1.0 / N, grad, loss, ws)
    grads[path] = grad.cpu().numpy().astype(np.float64)
scale = np.abs(want).max()
print("N", N, "scale", scale)
off = 0
for k, n in zip(keys, [int(np.prod(oppo.param_shapes(cont, O, A)[k])) for k in keys]):
    print(f"{k:28s} |want|max {np.abs(want[off:off+n]).max()/scale:9.2e}  " + "  ".join(f"{p} {np.abs(v[off:off+n]-want[off:off+n]).max()/scale:9.2e}" for p, v in grads.items()))
    off += n
# ---- forward accuracy: mu / std of the device forward (tiled fp32) and of torch-float32 against the float64 oracle
_, dist = ops.policy_act(params, True, O, A, 2.0, s, 1, 0, want_dist=True)
dist = dist.cpu().numpy().astype(np.float64)
for dtype, name in ((t.float32, "torch32"),):
    p = oppo.unflatten(g["init_flat"], cont, O, A)
    f64 = oppo.dist_params({k: v.double() for k, v in p.items()}, cont, t.from_numpy(r["states"]).double())[1]
    f32 = oppo.dist_params(p, cont, t.from_numpy(r["states"]))[1]
    mu64, sd64 = f64[0].numpy()[:, 0], f64[1].numpy()[:, 0]
    print("mu |max|", np.abs(mu64).max(), "std range", sd64.min(), sd64.max())
    print("abs err mu : device %.2e  torch32 %.2e" % (np.abs(dist[:, 0] - mu64).max(), np.abs(f32[0].numpy()[:, 0] - mu64).max()))
    print("abs err std: device %.2e  torch32 %.2e" % (np.abs(dist[:, 1] - sd64).max(), np.abs(f32[1].numpy()[:, 0] - sd64).max()))
    print("mean signed err mu: device %.2e torch32 %.2e" % ((dist[:, 0] - mu64).mean(), (f32[0].numpy()[:, 0] - mu64).mean()))
