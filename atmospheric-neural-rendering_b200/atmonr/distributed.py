"""Data-parallel ray sharding (new in this build; the reference is single-GPU,
scripts/train.py:94).

One process per GPU (torchrun), every rank holds the full parameters and ray tables, each global
batch is split into equal contiguous ray shards, every rank computes the MEAN loss of its shard
and the gradients are summed over NCCL and scaled by 1/world_size inside the fused AdamW kernel --
identical to the single-GPU global-mean loss when the shards are equal-sized. The small tensors
(MLPs) are all-reduced; the two hash tables are reduce-scattered, updated slice-wise and their fp16
shadow all-gathered (atmonr.optim.FusedAdamW.shard_large_parameters). Extraction shards voxel
columns with no communication.
"""

from __future__ import annotations

import os

import torch
import torch.distributed as td


def is_active() -> bool:
    return td.is_available() and td.is_initialized() and td.get_world_size() > 1


def rank() -> int:
    return td.get_rank() if (td.is_available() and td.is_initialized()) else 0


def world_size() -> int:
    return td.get_world_size() if (td.is_available() and td.is_initialized()) else 1


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment. Returns (rank, world, local_rank)."""
    if "RANK" not in os.environ or int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return 0, 1, 0
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not td.is_initialized():
        td.init_process_group(backend=backend)
    return td.get_rank(), td.get_world_size(), local


def shard_slice(n: int, rank_: int, world: int) -> slice:
    """Contiguous equal shards (the last one may be shorter)."""
    per = -(-n // world)
    return slice(min(rank_ * per, n), min((rank_ + 1) * per, n))


def all_reduce_gradients(optimizer) -> None:
    """Sum gradients over ranks; the 1/world_size factor is applied by FusedAdamW (grad_scale)
    or, for stock optimizers, here."""
    if not is_active():
        return
    world = td.get_world_size()
    fused = hasattr(optimizer, "grad_scale")
    if fused:
        optimizer.grad_scale = 1.0 / world
    sharded = getattr(optimizer, "is_sharded", lambda p: False)
    for group in optimizer.param_groups:
        for p in group["params"]:
            if p.grad is None or sharded(p):      # sharded tensors are reduce-scattered inside step()
                continue
            td.all_reduce(p.grad, op=td.ReduceOp.SUM)
            if not fused:
                p.grad.div_(world)


def broadcast_parameters(tensors) -> None:
    """Rank 0's values reach every rank. The broadcast writes in place under no_grad, which bumps the
    tensor's version counter, so a cached fp16 shadow (atmonr.native.modules.shadow_of) is rebuilt."""
    if not is_active():
        return
    with torch.no_grad():
        for t in tensors:
            td.broadcast(t, src=0)
            if hasattr(t, "_atmonr_shadow"):
                t._atmonr_shadow = None


def merge_disjoint_updates(values: torch.Tensor, touched: torch.Tensor) -> torch.Tensor:
    """values (..., n) holds this rank's updates at the columns flagged in touched (n,) bool; the ranks
    touched DISJOINT columns (ray shards of the same global batches). Returns values with every rank's
    updates merged in (columns nobody touched keep their local value, which is the same on all ranks).
    One SUM all-reduce of the masked values + one of the mask."""
    if not is_active():
        return values
    t = touched.to(values.dtype)
    upd = values * t
    td.all_reduce(upd, op=td.ReduceOp.SUM)
    td.all_reduce(t, op=td.ReduceOp.SUM)
    return torch.where(t > 0, upd, values)
