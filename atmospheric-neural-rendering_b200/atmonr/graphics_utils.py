"""Emission-absorption volume rendering (reference: src/atmonr/graphics_utils.py:6-77).

`render` / `render_with_surface` keep the reference's signatures and return values; the work is
one warp-per-ray kernel forward and one backward (atmonr_composite_fwd / _bwd), computed in
float32 whatever the input dtype (the reference inherits float16 from tiny-cuda-nn here).
`voxel_traversal` (graphics_utils.py:80-147) belongs to the globalgrid visualisation mode and
is out of this build's scope.
"""

from __future__ import annotations

import torch

from atmonr.native import ops


def render(z_vals: torch.Tensor, color: torch.Tensor, sigma: torch.Tensor):
    """-> (color_map (B,K), alpha (B,N,V), weights (B,N,V)). z_vals in km."""
    assert z_vals.dim() == 2 and color.dim() == 3 and sigma.dim() == 3
    assert z_vals.shape == color.shape[:2] and z_vals.shape == sigma.shape[:2]
    cmap, _, _, weights, alpha = ops.CompositeFn.apply(z_vals, color, sigma, None, 1.0, False)
    return cmap, alpha, weights


def render_with_surface(z_vals, color, sigma, color_surf):
    """-> (color_map, alpha, weights, color_map_atmo, color_map_surf); the surface sits behind
    the last sample with transmittance prod(1 - alpha)."""
    assert z_vals.dim() == 2 and color.dim() == 3 and sigma.dim() == 3
    cmap, catmo, csurf, weights, alpha = ops.CompositeFn.apply(z_vals, color, sigma, color_surf, 1.0, False)
    return cmap, alpha, weights, catmo, csurf


def voxel_traversal(*args, **kwargs):
    raise NotImplementedError("voxel_traversal (globalgrid extract mode) is outside the B200 hot-path build")
