"""ORACLE (test infrastructure, not product code): torch-CPU fp32 restatement of the reference's
ActorCritic / RND / PPO numerics, written functionally over a `state_dict`-keyed parameter dict.

Follows (file:line under /root/reference):
  ActorCritic forward pieces    PPO/ActorCritic.py:19-60   (Linear(no bias) -> GroupNorm(8,64) -> SiLU ...)
  get_dist / get_evaluate       PPO/ActorCritic.py:85-146
  PPO.compute_gae               PPO/PPO.py:107-120         (float32, NEP-50 promotion)
  PPO.learn                     PPO/PPO.py:122-260
  RND intrinsic / update_pred   PPO/RND.py:71-115
Pinned against the real reference by tests/golden/learn_*.npz (tests/test_oracle.py).

The distributions are written in closed form (what torch.distributions.Categorical(probs) and
MultivariateNormal(mu, diag(std^2)) compute), autograd supplies the gradients, and the optimiser is
a hand-written AdamW (lr, betas (0.9, 0.999), eps 1e-8, weight_decay 0.01 - torch defaults).
"""
from __future__ import annotations

import math

import numpy as np
import torch as t
import torch.nn.functional as F

H = 64
GROUPS = 8
GN_EPS = 1e-5
F32_EPS = float(np.finfo(np.float32).eps)


# ----------------------------------------------------------------------------------------- layout
def head_names(is_continuous: bool):
    return ["mu_head", "log_std_head", "critic"] if is_continuous else ["actor", "critic"]


def param_keys(is_continuous: bool):
    """state_dict / .parameters() order of the reference ActorCritic (ActorCritic.py:19-60)."""
    keys = ["model.0.weight", "model.1.weight", "model.1.bias"]
    for h in head_names(is_continuous):
        keys += [f"{h}.0.weight", f"{h}.1.weight", f"{h}.1.bias", f"{h}.3.weight", f"{h}.3.bias"]
    return keys


def param_shapes(is_continuous: bool, O: int, A: int):
    shp = {"model.0.weight": (H, O), "model.1.weight": (H,), "model.1.bias": (H,)}
    for h in head_names(is_continuous):
        out = 1 if h == "critic" else A
        shp.update({f"{h}.0.weight": (H, H), f"{h}.1.weight": (H,), f"{h}.1.bias": (H,),
                    f"{h}.3.weight": (out, H), f"{h}.3.bias": (out,)})
    return shp


def unflatten(flat, is_continuous, O, A):
    flat = t.as_tensor(flat, dtype=t.float32)
    out, off = {}, 0
    shp = param_shapes(is_continuous, O, A)
    for k in param_keys(is_continuous):
        n = int(np.prod(shp[k]))
        out[k] = flat[off:off + n].reshape(shp[k]).clone()
        off += n
    assert off == flat.numel()
    return out


def flatten(params, is_continuous):
    return t.cat([params[k].reshape(-1) for k in param_keys(is_continuous)])


RND_KEYS = ["0.weight", "0.bias", "1.weight", "1.bias", "3.weight", "3.bias"]


# ----------------------------------------------------------------------------------------- forward
def _block(x, w, gw, gb, b=None):
    return F.silu(F.group_norm(F.linear(x, w, b), GROUPS, gw, gb, GN_EPS))


def trunk(p, x):
    return _block(x, p["model.0.weight"], p["model.1.weight"], p["model.1.bias"])


def head(p, name, f):
    h = _block(f, p[f"{name}.0.weight"], p[f"{name}.1.weight"], p[f"{name}.1.bias"])
    return F.linear(h, p[f"{name}.3.weight"], p[f"{name}.3.bias"])


def dist_params(p, is_continuous, states):
    f = trunk(p, states)
    if is_continuous:
        mu = head(p, "mu_head", f)
        std = F.softplus(t.clamp(head(p, "log_std_head", f), -2, 2))
        return f, (mu, std)
    probs = t.softmax(head(p, "actor", f), dim=-1)
    return f, (probs,)


def evaluate(p, is_continuous, states, actions):
    """-> log_prob [b], value [b], mean entropy (scalar, no grad) - ActorCritic.get_evaluate."""
    f, dp = dist_params(p, is_continuous, states)
    if is_continuous:
        mu, std = dp
        k = mu.shape[-1]
        # MultivariateNormal(mu, diag(std^2)): scale_tril = cholesky = diag(sqrt(std^2))
        tril = t.sqrt(std * std)
        z = (actions - mu) / tril
        half_log_det = tril.log().sum(-1)
        logp = -0.5 * (k * math.log(2 * math.pi) + (z * z).sum(-1)) - half_log_det
        ent = 0.5 * k * (1.0 + math.log(2 * math.pi)) + half_log_det
    else:
        (probs,) = dp
        pn = probs / probs.sum(-1, keepdim=True)
        logits = t.log(t.clamp(pn, F32_EPS, 1 - F32_EPS))
        logp = logits.gather(-1, actions.long().reshape(-1, 1)).squeeze(-1)
        ent = -(logits * pn).sum(-1)
    value = head(p, "critic", f).squeeze(-1)
    return logp, value, ent.mean().detach()


# ----------------------------------------------------------------------------------------- GAE
def compute_gae(rewards, dones, values, next_value, gamma, gae_lambda):
    """numpy float32 scalar loop with the reference's promotion rules and evaluation order."""
    r = np.asarray(rewards, np.float32); d = np.asarray(dones, np.float32); v = np.asarray(values, np.float32)
    nv = np.float32(next_value)
    g = np.float32(gamma); gl = np.float32(gamma * gae_lambda)
    gae = np.float32(0)
    out = np.empty_like(v)
    one = np.float32(1)
    for i in range(len(v) - 1, -1, -1):
        nd = one - d[i]
        delta = r[i] + g * nv * nd - v[i]
        gae = delta + gl * nd * gae
        out[i] = gae + v[i]
        nv = v[i]
    return out


# ----------------------------------------------------------------------------------------- AdamW
class AdamW:
    def __init__(self, params, lr, wd=0.01, b1=0.9, b2=0.999, eps=1e-8):
        self.params, self.lr, self.wd, self.b1, self.b2, self.eps = params, lr, wd, b1, b2, eps
        self.m = [t.zeros_like(p) for p in params]
        self.v = [t.zeros_like(p) for p in params]
        self.step_count = 0

    @t.no_grad()
    def step(self, grads):
        self.step_count += 1
        bc1 = 1 - self.b1 ** self.step_count
        bc2 = 1 - self.b2 ** self.step_count
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            p.mul_(1 - self.lr * self.wd)
            m.mul_(self.b1).add_(g, alpha=1 - self.b1)
            v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(m, denom, value=-self.lr / bc1)


def clip_grad_norm(grads, max_norm):
    total = t.sqrt(sum((g.double() ** 2).sum() for g in grads)).to(grads[0].dtype)
    coef = t.clamp(max_norm / (total + 1e-6), max=1.0)
    return [g * coef for g in grads], total


# ----------------------------------------------------------------------------------------- RND
def rnd_net(p, prefix, x):
    h = _block(x, p[f"{prefix}.0.weight"], p[f"{prefix}.1.weight"], p[f"{prefix}.1.bias"], p[f"{prefix}.0.bias"])
    return F.linear(h, p[f"{prefix}.3.weight"], p[f"{prefix}.3.bias"])


def rnd_intrinsic(p, states, beta):
    with t.no_grad():
        return t.norm(rnd_net(p, "pred_net", states) - rnd_net(p, "target_net", states), dim=-1) * beta


def rnd_update(p, opt, states, mini_batch_size):
    keys = [f"pred_net.{k}" for k in RND_KEYS]
    for i in range(0, len(states), mini_batch_size):
        x = states[i:i + mini_batch_size]
        for k in keys:
            p[k].requires_grad_(True)
        loss = F.mse_loss(rnd_net(p, "pred_net", x), rnd_net(p, "target_net", x).detach())
        grads = t.autograd.grad(loss, [p[k] for k in keys])
        for k in keys:
            p[k].requires_grad_(False)
        opt.step(list(grads))


# ----------------------------------------------------------------------------------------- learn
def ppo_loss(p, is_continuous, s, a, old_logp, adv, ret, clip):
    logp, value, ent = evaluate(p, is_continuous, s, a)
    ratio = t.exp(t.clamp(logp - old_logp, -20, 20))
    surr1 = ratio * adv
    surr2 = t.clamp(ratio, 1 - clip, 1 + clip) * adv
    loss = -t.min(surr1, surr2) + 0.5 * F.smooth_l1_loss(value, ret) - 0.01 * ent
    return loss.mean()


def learn(params, is_continuous, mem, *, lr, k_epochs, policy_clip, gae_lambda, gamma, mini_batch_size,
          rnd_params=None, beta=0.001, opt=None, rnd_opt=None, trace=None, schedule=None, dtype=t.float32):
    """One PPO.learn() over `mem` = dict(states [N,O], actions, rewards [N], dones [N]) float32.
    Mutates `params` (and rnd_params) in place; returns the per-minibatch mean losses.

    `schedule` (optional): list of index arrays, one per minibatch, replacing the reference's sequential chunks
    `[k*mb, (k+1)*mb)` in the update loop - "the reference fed the same permutation" of SURVEY.md H7, used to check
    env-sharded runs whose global minibatch k is the union of every rank's k-th local chunk.  Old-policy evaluation,
    GAE and the advantage normalisation do not depend on it (row-wise / segment-wise / global).
    `dtype=torch.float64` evaluates the same update in double precision (the truth the float32 runs are judged by;
    GAE stays the reference's float32 loop); `params` must then be float64 tensors."""
    keys = param_keys(is_continuous)
    s = t.as_tensor(mem["states"], dtype=t.float32).to(dtype); a = t.as_tensor(mem["actions"], dtype=t.float32).to(dtype)
    N, mb = len(s), mini_batch_size
    old = {k: v.clone() for k, v in params.items()}
    with t.no_grad():
        lp, vals = [], []
        for i in range(0, N, mb):
            l, v, _ = evaluate(old, is_continuous, s[i:i + mb], a[i:i + mb])
            lp.append(l); vals.append(v)
        old_logp, old_values = t.cat(lp), t.cat(vals)
    rewards = np.asarray(mem["rewards"], np.float32)
    if rnd_params is not None:
        intr = t.cat([rnd_intrinsic(rnd_params, s[i:i + mb], beta) for i in range(0, N, mb)]).numpy()
        rewards = np.add(rewards, intr)
        if rnd_opt is None:
            rnd_opt = AdamW([rnd_params[f"pred_net.{k}"] for k in RND_KEYS], lr=1e-3)
        rnd_update(rnd_params, rnd_opt, s, mb)
    v_np = old_values.float().numpy()
    returns = t.from_numpy(compute_gae(rewards, mem["dones"], v_np, v_np[-1], gamma, gae_lambda)).to(dtype)
    adv = returns - old_values
    adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    if trace is not None:
        trace.update(old_logp=old_logp.numpy(), old_values=v_np, returns=returns.numpy(), advantages=adv.numpy(),
                     rewards=rewards)
    if opt is None:
        opt = AdamW([params[k] for k in keys], lr=lr)
    losses = []
    chunks = [slice(i, i + mb) for i in range(0, N, mb)] if schedule is None else [t.as_tensor(ix, dtype=t.long) for ix in schedule]
    for _ in range(k_epochs):
        for sl in chunks:
            for k in keys:
                params[k].requires_grad_(True)
            loss = ppo_loss(params, is_continuous, s[sl], a[sl], old_logp[sl], adv[sl], returns[sl], policy_clip)
            grads = t.autograd.grad(loss, [params[k] for k in keys])
            for k in keys:
                params[k].requires_grad_(False)
            grads, _ = clip_grad_norm(list(grads), 2.0)
            opt.step(grads)
            losses.append(float(loss))
    return losses
