"""clip_grad_norm_ + AdamW as ONE kernel over the flat parameter buffer (csrc/update.cu::k_adamw).

Reference: `optim.AdamW(policy.parameters(), lr)` with torch defaults betas (0.9, 0.999), eps 1e-8, weight_decay 0.01
(/root/reference/PPO/PPO.py:53-56) and `nn.utils.clip_grad_norm_(policy.parameters(), 2.0)` (:250); RND's optimiser at
/root/reference/PPO/RND.py:46-49 (lr 1e-3, no clipping).
"""
from __future__ import annotations

import torch

from . import ops


class FusedAdamW:
    def __init__(self, flat_params: torch.Tensor, lr: float, weight_decay: float = 0.01, max_norm: float = 2.0):
        self.params = flat_params
        self.lr, self.weight_decay, self.max_norm = float(lr), float(weight_decay), float(max_norm)
        self.exp_avg = torch.zeros_like(flat_params)
        self.exp_avg_sq = torch.zeros_like(flat_params)
        self.step_count = 0
        # the clock the kernel reads and advances: {int64 step, float64 beta1^step, float64 beta2^step} (zeros = step 0)
        self.step_dev = torch.zeros(3, dtype=torch.int64, device=flat_params.device)
        self.grad_norm = torch.zeros(1, dtype=torch.float64, device=flat_params.device)  # last pre-clip norm

    def step(self, grad: torch.Tensor) -> None:
        """One optimiser step.  The step number lives in device memory (`step_dev`) so the launch is replayable inside
        a CUDA graph; `step_count` mirrors it on the host for launches made directly."""
        self.step_count += 1
        ops.adamw_step_dev(self.params, grad, self.exp_avg, self.exp_avg_sq, self.step_dev, self.lr, self.weight_decay,
                           self.max_norm, self.grad_norm)

    def zero_grad(self, set_to_none: bool = True) -> None:  # API compatibility: gradients are produced whole by the kernels
        pass

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "lr": self.lr, "weight_decay": self.weight_decay, "max_norm": self.max_norm}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.step_dev.zero_()
        self.step_dev[0] = self.step_count   # the kernel recomputes the powers when it finds them zero
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
