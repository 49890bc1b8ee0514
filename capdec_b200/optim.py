"""AdamW with the semantics of `transformers.AdamW` 4.24 (what train.py:6,326 imports) on the fused sm_100a kernel,
plus the linear warm-up schedule (train.py:328-330).  Host code is plumbing: lr / step live in device scalars."""
from __future__ import annotations

import torch

from . import ops


class AdamW(torch.optim.Optimizer):
    """Drop-in for `transformers.AdamW(params, lr=...)`: betas (0.9, 0.999), eps 1e-6 added to sqrt(v) BEFORE bias
    correction, weight_decay 0 and decoupled.  One fused kernel launch per parameter tensor."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True):
        if not correct_bias:
            raise ValueError("correct_bias=False is not implemented (the reference uses the default True)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["lr_dev"] = torch.zeros(1, device=p.device)
                    st["t_dev"] = torch.zeros(1, device=p.device)
                st["step"] += 1
                st["lr_dev"].fill_(float(group["lr"]))
                st["t_dev"].fill_(float(st["step"]))
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                ops.adamw_step(p.data, g, st["exp_avg"], st["exp_avg_sq"], st["lr_dev"], st["t_dev"], b1, b2, group["eps"],
                               group["weight_decay"])
        return loss


def get_linear_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps, last_epoch=-1):
    """HF:optimization.py:101-104 (same LambdaLR the reference builds at train.py:328-330)."""

    def lr_lambda(current_step: int):
        if current_step < num_warmup_steps:
            return float(current_step) / float(max(1, num_warmup_steps))
        return max(0.0, float(num_training_steps - current_step) / float(max(1, num_training_steps - num_warmup_steps)))

    return torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda, last_epoch)
