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


def voxel_traversal(u: torch.Tensor, end: torch.Tensor, unique_only: bool = True) -> torch.Tensor:
    """Voxels (unit grid, any dimension) crossed by the segments u[i] -> end[i]; graphics_utils.py:80-147
    (Amanatides & Woo 1987), used by the global-grid extract layout (harp2_extract.py:858). Returns
    int16 voxel indices, one row per visit (all start voxels first, then one row per step), or the
    distinct rows when `unique_only`.

    A batched DDA on whatever device the inputs live on: every segment still under way advances along
    the axis whose next cell boundary is nearest; an axis that has reached the end voxel's index is
    frozen (its boundary distance becomes infinite), so a walk ends exactly in the end voxel, or stops
    as soon as some axis is past it. Plain torch: this runs once per extraction, not per step."""
    if u.shape != end.shape or u.dim() != 2:
        raise ValueError("u and end must both have shape (N, D)")
    delta = end - u
    direction = delta / torch.linalg.norm(delta, dim=-1, keepdim=True)
    step = torch.sign(direction).to(torch.int16)
    cell = torch.floor(u).to(torch.int16)
    last = torch.floor(end).to(torch.int16)
    # distance along the segment to the first boundary on each axis, and between boundaries
    along = step.to(u.dtype) * u
    t_next = torch.abs((torch.ceil(along) - along) / direction)
    t_next = torch.where(t_next.isnan() | (cell == last), torch.full_like(t_next, torch.inf), t_next)
    t_cell = torch.abs(1 / direction)
    visits = [torch.unique(cell, dim=0, sorted=False)]
    gap = (cell - last) * step                       # < 0: still short of the end index on that axis
    walking = ~((gap == 0).all(dim=-1) | (gap > 0).any(dim=-1))
    while bool(walking.any()):
        rows = torch.nonzero(walking)[:, 0]
        axis = torch.argmin(t_next[rows], dim=-1)
        t_next[rows, axis] += t_cell[rows, axis]
        cell[rows, axis] += step[rows, axis]
        visits.append(cell[rows].clone())
        gap = (cell[rows] - last[rows]) * step[rows]
        reached = gap >= 0
        t_next[rows] = torch.where(reached, torch.full_like(t_next[rows], torch.inf), t_next[rows])
        walking[rows] = ~(reached.all(dim=-1) | (gap > 0).any(dim=-1))
    out = torch.cat(visits, dim=0)
    return torch.unique(out, dim=0, sorted=False) if unique_only else out
