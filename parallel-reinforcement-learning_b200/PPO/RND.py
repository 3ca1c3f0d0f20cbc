"""Drop-in `RND` (reference: /root/reference/PPO/RND.py:8-115): frozen target net + trainable predictor net,
intrinsic reward = beta * ||pred(s) - target(s)||_2, predictor trained with MSE + AdamW(lr 1e-3).

Same constructor, `state_dict()` keys (`target_net.*`, `pred_net.*`) and initialisation as the reference; each net's
parameters are views into one flat float32 CUDA buffer that the fused kernels (csrc/update.cu: k_rnd_intrinsic,
k_rnd_grad, k_adamw) read and update in place.
"""
from __future__ import annotations

from copy import deepcopy

import torch as t
from torch import nn

from prl_b200 import ops
from prl_b200._lib import require_cuda
from prl_b200.optim import FusedAdamW

from .ActorCritic import flatten_module_


class RND(nn.Module):
    def __init__(self, in_features: int, out_features: int, beta: int = 0.001):
        super().__init__()
        self.device = require_cuda()
        self.in_features, self.out_features = int(in_features), int(out_features)

        model = nn.Sequential(nn.Linear(in_features, 64), nn.GroupNorm(64 // 8, 64), nn.SiLU(inplace=True),
                              nn.Linear(64, out_features))
        # RND.py:33-38: both nets are deep copies made BEFORE init_weights(), so they get different random weights
        self.target_net = deepcopy(model)
        self.pred_net = deepcopy(model)
        self.init_weights()
        for param in self.target_net.parameters():
            param.requires_grad = False

        self.beta = beta
        self.loss_fn = nn.MSELoss()
        self.target_flat = flatten_module_(self.target_net, self.device)
        self.pred_flat = flatten_module_(self.pred_net, self.device)
        assert self.pred_flat.numel() == ops.rnd_param_count(self.in_features, self.out_features)
        self.optimizer = FusedAdamW(self.pred_flat, lr=0.001, max_norm=0.0)  # RND.py:46-49; no gradient clipping
        self._grad = t.zeros_like(self.pred_flat)
        self._loss = t.zeros(4, dtype=t.float64, device=self.device)
        self._ws = None
        self.eval()

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.normal_(m.bias, mean=0, std=0.01)
            elif isinstance(m, nn.GroupNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def _apply(self, fn, recurse=True):
        if getattr(self, "pred_flat", None) is not None:
            raise RuntimeError("RND parameters live in flat CUDA buffers; .to()/.cpu()/.half() are not supported")
        return super()._apply(fn, recurse)

    def _chunk(self, x: t.Tensor) -> t.Tensor:
        return x.to(device=self.device, dtype=t.float32).reshape(-1, self.in_features).contiguous()

    @t.no_grad()
    def intrinsic_reward_device(self, states: t.Tensor, add_to: t.Tensor | None = None, out: t.Tensor | None = None):
        """One launch over all rows: beta * ||pred - target||_2 (+ add_to).  Row-independent, so chunking is moot."""
        return ops.rnd_intrinsic(self.target_flat, self.pred_flat, self.in_features, self.out_features, states, self.beta,
                                 add_to=add_to, out=out)

    @t.no_grad()
    def compute_intrinsic_reward(self, values) -> t.Tensor:
        """RND.py:71-94.  `values`: any iterable of [b, in_features] tensors (list from batch_packer, DataLoader...)."""
        return t.cat([self.intrinsic_reward_device(self._chunk(v)) for v in values], dim=0)

    @t.no_grad()
    def update_pred_chunk(self, states: t.Tensor, comm=None, global_rows: int | None = None):
        """One MSE(mean) + AdamW step on `states`.  Sharded (`comm`): this rank's rows are its part of a global chunk of
        `global_rows` rows - the local mean-gradient is re-weighted by n / global_rows and summed over the ranks (one
        allreduce of the flat predictor gradient), so every rank applies the same step to its replica."""
        n = states.shape[0]
        if n > 0:
            need = ops.update_ws_floats(False, self.in_features, self.out_features, n) + 4096
            if self._ws is None or self._ws.numel() < need:
                self._ws = t.empty(need, dtype=t.float32, device=self.device)
            ops.rnd_grad(self.target_flat, self.pred_flat, self.in_features, self.out_features, states, self._grad, self._loss, self._ws)
        else:
            self._grad.zero_()
        if comm is not None:
            self._grad.mul_(n / float(global_rows))
            comm.allreduce_(self._grad)
        self.optimizer.step(self._grad)

    def update_pred(self, values) -> None:
        """RND.py:96-115: one pass, one MSE(mean) + AdamW step per chunk, in order."""
        for v in values:
            self.update_pred_chunk(self._chunk(v))
