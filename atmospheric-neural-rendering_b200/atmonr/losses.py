"""Per-band losses (reference: src/atmonr/losses.py). Same names and signatures; on CUDA
tensors each is one fused kernel producing the loss and its gradient
(atmonr_band_loss, csrc/atmonr_b200.cu)."""

from __future__ import annotations

import torch

from atmonr.native import ops


def _apply(kind: str, pred: torch.Tensor, gt: torch.Tensor, max_i: float) -> torch.Tensor:
    band = torch.zeros(pred.shape[0], dtype=torch.int64, device=pred.device)
    return ops.band_loss(pred.reshape(-1, 1), band, gt.reshape(-1), max_i, kind)


def dark_loss(pred, gt, max_i):
    return _apply("dark", pred, gt, max_i)


def hdr_loss(pred, gt, max_i):
    return _apply("hdr", pred, gt, max_i)


def l1_loss(pred, gt, max_i):
    return _apply("l1", pred, gt, max_i)


def l1_plus_hdr_loss(pred, gt, max_i):
    return _apply("l1_plus_hdr", pred, gt, max_i)


def mse_loss(pred, gt, max_i):
    return _apply("mse", pred, gt, max_i)


def mse_plus_hdr_loss(pred, gt, max_i):
    return _apply("mse_plus_hdr", pred, gt, max_i)
