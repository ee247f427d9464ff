"""Spherical-Earth helpers (reference: src/atmonr/geospatial/spherical.py). Only used by the
extraction / visualisation callers; EARTH_RADIUS is imported by scripts/extract.py."""

from __future__ import annotations

import torch

from atmonr.geospatial.wgs_84 import WGS_84_A, WGS_84_B

EARTH_RADIUS = 6.378e6  # metres


def wgs_84_to_spherical(xyz: torch.Tensor) -> torch.Tensor:
    z = xyz[..., 2:] * WGS_84_A / WGS_84_B
    return torch.cat([xyz[..., :2], z], dim=-1) * EARTH_RADIUS / WGS_84_A


def spherical_to_wgs84(xyz: torch.Tensor) -> torch.Tensor:
    out = xyz * WGS_84_A / EARTH_RADIUS
    out[..., 2] *= WGS_84_B / WGS_84_A
    return out


def stretch_above_sea_level(xyz: torch.Tensor, stretch: float) -> torch.Tensor:
    r = torch.linalg.norm(xyz, dim=-1)
    factor = torch.where(r > EARTH_RADIUS, ((r - EARTH_RADIUS) * stretch + EARTH_RADIUS) / r, torch.ones_like(r))
    return xyz * factor[..., None]
