"""Drop-in `ActorCritic` (reference: /root/reference/PPO/ActorCritic.py:13-146).

Same constructor, module tree, `state_dict()` keys / shapes and Xavier initialisation as the reference, so checkpoints
move both ways - but every parameter is a VIEW into one flat float32 CUDA buffer (`self.flat`, laid out in
`.parameters()` order = csrc/mlp.cuh::PolicyLayout), which is what the sm_100a kernels read and the fused AdamW
kernel updates in place.  The nn.Module is a parameter container only: `forward` math never runs through torch here;
`get_dist / get_state_value / get_evaluate` launch the fused forward kernels through the C ABI.
"""
from __future__ import annotations

import torch as t
from torch import distributions, nn

from prl_b200 import ops
from prl_b200._lib import require_cuda


def _head(out_features: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(64, 64, bias=False), nn.GroupNorm(64 // 8, 64), nn.SiLU(inplace=True),
                         nn.Linear(64, out_features))


def flatten_module_(module: nn.Module, device) -> t.Tensor:
    """Move `module`'s parameters into ONE flat float32 buffer on `device` (parameters() order) and re-point every
    Parameter at its slice, so in-place kernel updates of the buffer are visible through state_dict()."""
    params = list(module.parameters())
    flat = t.empty(sum(p.numel() for p in params), dtype=t.float32, device=device)
    off = 0
    for p in params:
        n = p.numel()
        flat[off:off + n].copy_(p.detach().reshape(-1))
        p.data = flat[off:off + n].view(p.shape)
        off += n
    return flat


class ActorCritic(nn.Module):
    def __init__(self, is_continuous: bool, observ_dim: int, action_dim: int):
        super().__init__()
        self.device = require_cuda()
        self.is_continuous = is_continuous
        self.observ_dim, self.action_dim = int(observ_dim), int(action_dim)

        # ActorCritic.py:19-60 - same registration order, hence the same parameters()/state_dict order
        self.model = nn.Sequential(nn.Linear(observ_dim, 64, bias=False), nn.GroupNorm(64 // 8, 64), nn.SiLU(inplace=True))
        if self.is_continuous:
            self.mu_head = _head(action_dim)
            self.log_std_head = _head(action_dim)
        else:
            self.actor = _head(action_dim)
            self.actor.append(nn.Softmax(dim=-1))
        self.critic = _head(1)

        self.init_weights()
        self.flat = flatten_module_(self, self.device)
        assert self.flat.numel() == ops.policy_param_count(is_continuous, self.observ_dim, self.action_dim)

    def init_weights(self):
        """ActorCritic.py:66-80: Xavier-uniform Linear weights, N(0, 0.01) biases, GroupNorm weight 1 / bias 0."""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.normal_(m.bias, mean=0, std=0.01)
            elif isinstance(m, nn.GroupNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    # nn.Module.to()/cuda() would re-allocate the parameters and break the flat aliasing
    def _apply(self, fn, recurse=True):
        if getattr(self, "flat", None) is not None:
            raise RuntimeError("ActorCritic parameters live in one flat CUDA buffer; .to()/.cpu()/.half() are not supported")
        return super()._apply(fn, recurse)

    def forward(self, state: t.Tensor):
        raise NotImplementedError

    def _states(self, state: t.Tensor) -> t.Tensor:
        return state.to(device=self.device, dtype=t.float32).reshape(-1, self.observ_dim).contiguous()

    @t.no_grad()
    def dist_params(self, state: t.Tensor):
        """probs [n, A] (discrete) or (mu [n, A], std [n, A]) from the fused forward kernel."""
        s = self._states(state)
        _, dist = ops.policy_act(self.flat, self.is_continuous, self.observ_dim, self.action_dim, 1.0, s, seed=0,
                                 call_index=0, want_dist=True)
        if self.is_continuous:
            return dist[:, :self.action_dim], dist[:, self.action_dim:]
        return dist

    def get_dist(self, state: t.Tensor):
        """ActorCritic.py:85-110.  Returns a torch distribution built on the kernel's outputs (API compatibility; the
        hot path samples inside the kernels: PPO.get_action / AsyncPPO.worker)."""
        if self.is_continuous:
            mu, std = self.dist_params(state)
            return distributions.MultivariateNormal(mu, scale_tril=t.diag_embed(std))
        return distributions.Categorical(self.dist_params(state))

    @t.no_grad()
    def get_state_value(self, state: t.Tensor):
        s = self._states(state)
        dummy = t.zeros(s.shape[0], self.action_dim if self.is_continuous else 1, device=self.device)
        _, value, _ = ops.policy_evaluate(self.flat, self.is_continuous, self.observ_dim, self.action_dim, s, dummy)
        return value

    @t.no_grad()
    def get_evaluate(self, states: t.Tensor, actions: t.Tensor):
        """ActorCritic.py:118-146 -> (log_probs [b], state_value [b], mean entropy (0-dim))."""
        s = self._states(states)
        n = s.shape[0]
        a = actions.to(device=self.device, dtype=t.float32).reshape(n, -1)
        a = a[:, :self.action_dim] if self.is_continuous else a[:, :1]
        logp, value, ent = ops.policy_evaluate(self.flat, self.is_continuous, self.observ_dim, self.action_dim, s, a.contiguous())
        return logp, value, (ent / n).to(t.float32).squeeze(0)
