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


class _InverseCdfFn(torch.autograd.Function):
    """z of the fine samples (sorted together with the coarse ones). Forward is the
    atmonr_sample_pdf kernel; backward follows the reference's graph: only the bin width is
    detached (samplers.py:96), so gradients reach the coarse weights through the CDF and the
    coarse z through the concatenation."""

    @staticmethod
    def forward(ctx, weights, z_coarse, u):
        z, inds = ops.sample_pdf_z(weights, z_coarse, u)
        ctx.mark_non_differentiable(inds)
        return z, inds

    @staticmethod
    def backward(ctx, gz, _gi):
        raise NotImplementedError(
            "differentiating through sample_pdf: use atmonr.samplers.sample_pdf(differentiable=True)"
        )


def _inverse_cdf_torch(w, z_coarse, u):
    """Differentiable path (torch graph, same operations as samplers.py:72-101)."""
    w = w[:, 1:-1]
    pdf = (w + 1e-8) / torch.sum(w + 1e-8, dim=1, keepdim=True)
    cdf = torch.cumsum(pdf, dim=1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=1)
    inds = torch.searchsorted(cdf.detach(), u.contiguous(), right=True)
    lo = torch.clamp(inds - 1, min=0)
    hi = torch.clamp(inds, max=cdf.shape[-1] - 1)
    mids = 0.5 * (z_coarse[..., 1:] + z_coarse[..., :-1])
    c_lo, c_hi = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
    m_lo, m_hi = torch.gather(mids, 1, lo), torch.gather(mids, 1, hi)
    den = c_hi - c_lo
    den = torch.where(den < 1e-8, torch.ones_like(den), den)
    fine = m_lo + (u - c_lo) / den * (m_hi - m_lo).detach()
    return torch.sort(torch.cat([z_coarse, fine], -1), -1)[0]


def sample_pdf(ray_batch, pdf_discrete, z_vals_c, n_samples: int = 128):
    """samplers.py:50-103 -> pts (B, N_c+n_samples, 3), z_vals (B, N_c+n_samples)."""
    w = pdf_discrete[..., 0]
    u = torch.rand((w.shape[0], n_samples), device=w.device)
    if torch.is_grad_enabled() and (w.requires_grad or z_vals_c.requires_grad):
        z = _inverse_cdf_torch(w, z_vals_c, u)
    else:
        z, _ = _InverseCdfFn.apply(w, z_vals_c, u)
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
    """samplers.py:168-195: append ellipsoidal height / ray_origin_height as a 4th coordinate."""
    from atmonr.geospatial.wgs_84 import cartesian_to_horizontal

    xyz = pts.double() * scale + offset[None, None]
    alt = cartesian_to_horizontal(xyz[..., 0], xyz[..., 1], xyz[..., 2])[2]
    return torch.cat([pts, (alt / ray_origin_height).float()[..., None]], dim=-1)
