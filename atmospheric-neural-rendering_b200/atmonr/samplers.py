"""Ray samplers (reference: src/atmonr/samplers.py). Same names, arguments and return shapes;
CUDA tensors in, one kernel per call. The uniforms come from torch's generator on the batch's
device, exactly where the reference draws them, so seeding behaves identically."""

from __future__ import annotations

from typing import Mapping

import torch

from atmonr.native import ops


def sample_uniform_bins(ray_batch: Mapping[str, torch.Tensor], n_bins: int = 64, random: bool = True):
    """samplers.py:8-47 -> pts (B, n_bins, 3), z_vals (B, n_bins)."""
    origin = ray_batch["origin"]
    u = torch.rand((origin.shape[0], n_bins), device=origin.device) if random else None
    return ops.sample_uniform(origin, ray_batch["dir"], ray_batch["len"], n_bins, u=u, random=False)


def inverse_cdf_z(weights: torch.Tensor, z_vals_c: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """samplers.py:72-101: sorted union of the coarse distances and the inverse-CDF samples; differentiable
    like the reference's graph (kernels forward and backward, atmonr.native.ops.InverseCdfFn)."""
    return ops.InverseCdfFn.apply(weights, z_vals_c, u)[0]


def sample_pdf(ray_batch, pdf_discrete, z_vals_c, n_samples: int = 128):
    """samplers.py:50-103 -> pts (B, N_c+n_samples, 3), z_vals (B, N_c+n_samples)."""
    w = pdf_discrete[..., 0]
    u = torch.rand((w.shape[0], n_samples), device=w.device)
    z = inverse_cdf_z(w, z_vals_c, u)
    pts = ray_batch["origin"][:, None] + ray_batch["dir"][:, None] * z[..., None]
    return pts, z


def sample_biased_bins(ray_batch, n_bins: int, ray_origin_height: float, alpha: float):
    """samplers.py:106-165. Unused by both shipped pipelines; kept for API
    completeness as a direct closed-form evaluation."""
    assert 0 <= alpha <= 1
    origin = ray_batch["origin"]
    norm = (alpha + 1) / 2
    edges = torch.linspace(0, 1, n_bins + 1, device=origin.device)[None]
    flat = edges[:, :-1] + torch.rand((origin.shape[0], n_bins), device=origin.device) / n_bins
    if alpha == 1:
        z = flat.clone()
    else:
        z = (-alpha + torch.sqrt(alpha**2 + 2 * (1 - alpha) * norm * flat)) * (1 / (1 - alpha))
    z = torch.where(flat <= 1, z, torch.ones_like(z)) * ray_batch["len"][:, None]
    return origin[:, None] + ray_batch["dir"][:, None] * z[..., None], z


def append_heights(pts, ray_origin_height: float, scale: float, offset: torch.Tensor):
    """samplers.py:168-195: append ellipsoidal height / ray_origin_height as a 4th coordinate
    (atmonr_append_heights; the float64 torch expressions only where the points carry a gradient)."""
    if not (torch.is_grad_enabled() and pts.requires_grad) and pts.dtype == torch.float32:
        return ops.append_heights(pts, ray_origin_height, scale, offset.tolist())
    from atmonr.geospatial.wgs_84 import cartesian_to_horizontal

    xyz = pts.double() * scale + offset[None, None]
    alt = cartesian_to_horizontal(xyz[..., 0], xyz[..., 1], xyz[..., 2])[2]
    return torch.cat([pts, (alt / ray_origin_height).float()[..., None]], dim=-1)
