"""Oracle: emission-absorption compositing and the per-band losses.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Reference: src/atmonr/graphics_utils.py:6-77
and src/atmonr/losses.py.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F


def voronoi_deltas(z):
    """graphics_utils.py:30-35: cell boundaries are 0, the sample mid-points, and the LAST
    SAMPLE (not the ray end)."""
    mid = (z[..., :-1] + z[..., 1:]) / 2
    edges = torch.cat([z[..., :1] * 0, mid, z[..., -1:]], dim=-1)
    return torch.diff(edges, dim=-1)


def composite(z, color, sigma):
    """graphics_utils.py:6-49 (render).  z (B,N) in km, color (B,N,K), sigma (B,N,1|K)."""
    z = z.to(color.dtype)
    delta = voronoi_deltas(z)[..., None]
    alpha = 1 - torch.exp(-sigma * delta)
    lead = torch.ones_like(alpha[:, :1])
    trans = torch.cumprod(torch.cat([lead, 1 - alpha + 1e-10], dim=1), dim=1)[:, :-1]
    weights = alpha * trans
    return torch.sum(color * weights, dim=1), alpha, weights


def composite_with_surface(z, color, sigma, color_surf):
    """graphics_utils.py:52-77 (render_with_surface); the surface transmittance is the plain
    product of (1 - alpha), without the +1e-10 used inside the cumprod."""
    c_atmo, alpha, weights = composite(z, color, sigma)
    c_surf = (1 - alpha).prod(dim=1) * color_surf
    return c_atmo + c_surf, alpha, weights, c_atmo, c_surf


# ---- losses.py:5-33 ---------------------------------------------------------------------
def dark_loss(pred, gt, max_i):
    return (((pred - gt) / (pred.detach() + 1e-3 * max_i)) ** 2).mean()


def hdr_loss(pred, gt, max_i):
    return F.mse_loss(torch.log(gt + 1e-3 * max_i), torch.log(pred + 1e-3 * max_i))


def l1_loss(pred, gt, max_i):
    return F.l1_loss(pred / max_i, gt / max_i)


def mse_loss(pred, gt, max_i):
    return F.mse_loss(pred / max_i, gt / max_i)


def l1_plus_hdr_loss(pred, gt, max_i):
    return l1_loss(pred, gt, max_i) + 0.2 * hdr_loss(pred, gt, max_i)


def mse_plus_hdr_loss(pred, gt, max_i):
    return mse_loss(pred, gt, max_i) + 0.2 * hdr_loss(pred, gt, max_i)


LOSSES = {
    "dark": dark_loss,
    "hdr": hdr_loss,
    "l1": l1_loss,
    "l1_plus_hdr": l1_plus_hdr_loss,
    "mse": mse_loss,
    "mse_plus_hdr": mse_plus_hdr_loss,
}


def band_select(color_map, irgb_idx):
    """instant_ngp.py:259-261 / nerf.py:232-237."""
    return torch.take_along_dim(color_map, irgb_idx[:, None], 1)[:, 0]
