"""FusedAdamW: torch.optim.AdamW semantics in one kernel per parameter tensor.

Mirrors the arithmetic of torch's single-tensor AdamW (decoupled decay, lerp first moment,
addcmul second moment, bias corrections combined in double precision on the host). It is a real
torch.optim.Optimizer: param_groups / state_dict have AdamW's layout (`step`, `exp_avg`,
`exp_avg_sq`), so ExponentialLR and the Trainer's checkpointing work unchanged
(reference: instant_ngp.py:107-127, trainer.py:53-67,239-274).

In the same pass the kernel refreshes the fp16 shadow copy the hash-grid / MLP kernels read,
and (optionally) divides the gradient by `grad_scale` and zeroes it.

Data-parallel runs: `shard_large_parameters()` switches the large tensors (the two hash tables:
42.3 M + 5.5 M parameters of 47.8 M) from "all-reduce the gradient, every rank updates everything" to
    reduce-scatter(gradient) -> AdamW on this rank's 1/G slice -> all-gather(fp16 shadow slice):
the same bytes leave each GPU over NVLink as a ring all-reduce moves (a reduce-scatter plus an all-gather of
HALF-width values: 3/4 of it), the optimizer's HBM traffic drops by G, and the kernels only ever read the
fp16 shadow, which is complete on every rank after the all-gather. The float32 master copy and the two
moments of the OTHER ranks' slices go stale in between and are brought up to date by `consolidate()`
(all-gather of the three float32 tensors; the Trainer calls it on every rank before a checkpoint is
written), so `state_dict()` keeps AdamW's full-size layout and checkpoints interchange.
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
        self.shard_min_numel = 0   # > 0: parameters at least this large are updated slice-wise (see above)
        self._grad_slices: dict = {}

    # ------------------------------------------------------------------ data-parallel sharding
    def shard_large_parameters(self, min_numel: int = 1 << 20) -> None:
        self.shard_min_numel = int(min_numel)

    def is_sharded(self, p: torch.Tensor) -> bool:
        import torch.distributed as td
        if self.shard_min_numel <= 0 or not (td.is_available() and td.is_initialized()):
            return False
        w = td.get_world_size()
        return w > 1 and p.numel() >= self.shard_min_numel and p.numel() % (w * 8) == 0

    def _slice(self, p: torch.Tensor):
        import torch.distributed as td
        n = p.numel() // td.get_world_size()
        lo = td.get_rank() * n
        return lo, lo + n

    def _reduce_scatter(self, p: torch.Tensor) -> torch.Tensor:
        """Sum of every rank's gradient over this rank's slice (float32, persistent buffer)."""
        import torch.distributed as td
        lo, hi = self._slice(p)
        buf = self._grad_slices.get(p)
        if buf is None or buf.numel() != hi - lo or buf.device != p.device:
            buf = self._grad_slices[p] = torch.empty(hi - lo, device=p.device, dtype=p.dtype)
        flat = p.grad.reshape(-1)
        if td.get_backend() == "nccl":
            td.reduce_scatter_tensor(buf, flat, op=td.ReduceOp.SUM)
        else:   # gloo (CPU test suite) has no reduce-scatter: all-reduce and keep the slice
            td.all_reduce(flat, op=td.ReduceOp.SUM)
            buf.copy_(flat[lo:hi])
        return buf

    @torch.no_grad()
    def consolidate(self) -> None:
        """Every rank's float32 master slices and moment slices -> every rank (collective: call it on ALL
        ranks, e.g. before rank 0 writes a checkpoint). No-op unless parameters are sharded."""
        import torch.distributed as td
        for group in self.param_groups:
            for p in group["params"]:
                if not self.is_sharded(p) or p not in self.state or not self.state[p]:
                    continue
                lo, hi = self._slice(p)
                for t in (p.data.reshape(-1), self.state[p]["exp_avg"].reshape(-1), self.state[p]["exp_avg_sq"].reshape(-1)):
                    td.all_gather_into_tensor(t, t[lo:hi].clone())

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
                if self.is_sharded(p):
                    self._step_slice(p, state, group)
                    continue
                grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                # a persistent gradient buffer of the fused backward: zeroed by the kernel that consumes it
                flag = PERSISTENT_GRADS.get(grad.data_ptr())
                ops.adamw_step(p.data, grad, state["exp_avg"], state["exp_avg_sq"], shadow_of(p),
                               group["lr"], beta1, beta2, group["eps"], group["weight_decay"],
                               int(state["step"].item()), grad_scale=self.grad_scale, zero_grad=flag is not None)
                if flag is not None:
                    flag[0] = True
        return loss

    def _step_slice(self, p, state, group) -> None:
        """reduce-scatter -> AdamW on this rank's slice -> all-gather of the fp16 shadow slice."""
        import torch.distributed as td
        beta1, beta2 = group["betas"]
        lo, hi = self._slice(p)
        g = self._reduce_scatter(p)
        flag = PERSISTENT_GRADS.get(p.grad.data_ptr())
        if flag is not None:          # the full-size buffer is reused by the next backward: clear it here
            p.grad.zero_()
            flag[0] = True
        shadow = shadow_of(p).reshape(-1)
        ops.adamw_step(p.data.reshape(-1)[lo:hi], g, state["exp_avg"].reshape(-1)[lo:hi],
                       state["exp_avg_sq"].reshape(-1)[lo:hi], shadow[lo:hi], group["lr"], beta1, beta2, group["eps"],
                       group["weight_decay"], int(state["step"].item()), grad_scale=1.0 / td.get_world_size())
        td.all_gather_into_tensor(shadow, shadow[lo:hi].clone() if td.get_backend() != "nccl" else shadow[lo:hi])
