"""FusedAdamW: torch.optim.AdamW semantics in one kernel per parameter tensor.

Mirrors the arithmetic of torch's single-tensor AdamW (decoupled decay, lerp first moment,
addcmul second moment, bias corrections combined in double precision on the host). It is a real
torch.optim.Optimizer: param_groups / state_dict have AdamW's layout (`step`, `exp_avg`,
`exp_avg_sq`), so ExponentialLR and the Trainer's checkpointing work unchanged
(reference: instant_ngp.py:107-127, trainer.py:53-67,239-274).

In the same pass the kernel refreshes the fp16 shadow copy the hash-grid / MLP kernels read,
and (optionally) divides the gradient by `grad_scale` and zeroes it.
"""

from __future__ import annotations

import torch
from torch.optim import Optimizer

from atmonr.native import ops
from atmonr.native.fused import PERSISTENT_GRADS
from atmonr.native.modules import shadow_of


class FusedAdamW(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.grad_scale = 1.0  # gradients are multiplied by this before use (1/world_size under DP)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None or p.numel() == 0:
                    continue
                state = self.state[p]
                if not state:
                    state["step"] = torch.tensor(0.0)
                    state["exp_avg"] = torch.zeros_like(p)
                    state["exp_avg_sq"] = torch.zeros_like(p)
                state["step"] += 1
                grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                # a persistent gradient buffer of the fused backward: zeroed by the kernel that consumes it
                flag = PERSISTENT_GRADS.get(grad.data_ptr())
                ops.adamw_step(p.data, grad, state["exp_avg"], state["exp_avg_sq"], shadow_of(p),
                               group["lr"], beta1, beta2, group["eps"], group["weight_decay"],
                               int(state["step"].item()), grad_scale=self.grad_scale, zero_grad=flag is not None)
                if flag is not None:
                    flag[0] = True
        return loss
