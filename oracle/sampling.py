"""Oracle: ray samplers.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference: src/atmonr/samplers.py.  The uniform draws are an explicit argument so the
oracle and the CUDA path can be fed identical random numbers.
"""

from __future__ import annotations

import torch

from oracle.geodesy import ecef_to_geodetic


def stratified_z(length, n_bins, u=None):
    """samplers.py:34-42.  u: (B, n_bins) uniforms in [0,1) or None for bin mid-points."""
    edges = torch.linspace(0, 1, n_bins + 1)[None]
    t = 0.5 if u is None else u
    return (edges[:, :-1] + t / n_bins) * length[:, None]


def points_on_rays(origin, direction, z):
    """samplers.py:45."""
    return origin[:, None] + direction[:, None] * z[..., None]


def sample_uniform(origin, direction, length, n_bins, u=None):
    """samplers.py:8-47 (sample_uniform_bins)."""
    z = stratified_z(length, n_bins, u)
    return points_on_rays(origin, direction, z), z


def inverse_cdf_z(weights, z_coarse, u):
    """samplers.py:72-101 (the body of sample_pdf up to the sort).

    weights: (B, N_c, 1); z_coarse: (B, N_c); u: (B, n_samples).
    Returns (z_sorted (B, N_c+n_samples), inds (B, n_samples) int64).  The gradient path is
    the reference's: only the bin *width* is detached (samplers.py:96)."""
    w = weights[:, 1:-1, 0]
    pdf = (w + 1e-8) / torch.sum(w + 1e-8, dim=1, keepdim=True)
    cdf = torch.cumsum(pdf, dim=1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=1)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    lo = torch.clamp(inds - 1, min=0)
    hi = torch.clamp(inds, max=cdf.shape[-1] - 1)
    mids = 0.5 * (z_coarse[..., 1:] + z_coarse[..., :-1])
    cdf_lo, cdf_hi = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
    z_lo, z_hi = torch.gather(mids, 1, lo), torch.gather(mids, 1, hi)
    den = cdf_hi - cdf_lo
    den = torch.where(den < 1e-8, torch.ones_like(den), den)
    t = (u - cdf_lo) / den
    fine = z_lo + t * (z_hi - z_lo).detach()
    z, _ = torch.sort(torch.cat([z_coarse, fine], -1), -1)
    return z, inds


def sample_pdf(origin, direction, weights, z_coarse, u):
    """samplers.py:50-103."""
    z, inds = inverse_cdf_z(weights, z_coarse, u)
    return points_on_rays(origin, direction, z), z, inds


def append_heights(pts, origin_height, scale, offset):
    """samplers.py:168-195."""
    xyz = pts.double() * scale + offset[None, None]
    alt = ecef_to_geodetic(xyz[..., 0], xyz[..., 1], xyz[..., 2])[2]
    return torch.cat([pts, (alt / origin_height).float()[..., None]], dim=-1)
