"""Oracle: WGS-84 geodesy, ray construction and the 'horizontal' point preprocessor.

TEST INFRASTRUCTURE (see oracle/__init__.py).  torch-CPU restatement; float64 where the
reference computes in float64.  Reference: src/atmonr/geospatial/wgs_84.py and
src/atmonr/datasets/harp2.py:351-390.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import torch

# wgs_84.py:17-21 -- note that the reference's "E" and "E2" are the *squared* first and
# second eccentricities.
A = 6378137.0
B = 6356752.314245
E_SQ = (A * A - B * B) / (A * A)
EP_SQ = (A * A - B * B) / (B * B)


def geodetic_to_ecef(lat_deg, lon_deg, alt):
    """wgs_84.py:24-53 (horizontal_to_cartesian).  Works in the dtype it is given."""
    phi = lat_deg * math.pi / 180
    lam = lon_deg * math.pi / 180
    s = torch.sin(phi)
    n = A / torch.sqrt(1 - (E_SQ * s**2))
    c = torch.cos(phi)
    x = (n + alt) * c * torch.cos(lam)
    y = (n + alt) * c * torch.sin(lam)
    z = (n * (1 - E_SQ) + alt) * s
    return x, y, z


def ecef_to_geodetic(x, y, z):
    """wgs_84.py:56-97 (cartesian_to_horizontal): ONE Bowring iteration, and the height is
    recovered as x / (cos(lat) cos(lon)) - N.  Restated literally (do not
    'improve' it).  Returns degrees, degrees, metres."""
    lam = torch.atan2(y, x)
    d = torch.sqrt(x**2 + y**2)
    u = torch.atan2(z / d, torch.zeros_like(x) + A / B)
    phi = torch.atan2(
        z + (EP_SQ * B) * (torch.sin(u) ** 3),
        d - (E_SQ * A) * (torch.cos(u) ** 3),
    )
    n = A / torch.sqrt(1 - (E_SQ * torch.sin(phi) ** 2))
    alt = x / (torch.cos(phi) * torch.cos(lam)) - n
    return phi * 180 / torch.pi, lam * 180 / torch.pi, alt


def _rot(theta_deg, phi_deg):
    """wgs_84.py:100-132: rotation matrices from (zenith, azimuth), angles negated."""
    th = -theta_deg * torch.pi / 180
    ph = -phi_deg * torch.pi / 180
    st, ct, sp, cp = torch.sin(th), torch.cos(th), torch.sin(ph), torch.cos(ph)
    zero = torch.zeros_like(th)
    rows = [
        torch.stack([cp, -sp * ct, sp * st], dim=1),
        torch.stack([sp, cp * ct, -cp * st], dim=1),
        torch.stack([zero, st, ct], dim=1),
    ]
    return torch.stack(rows, dim=1)


def view_dirs_local(theta_deg, phi_deg):
    """wgs_84.py:135-160: rotate +z by the (zenith, azimuth) matrix."""
    shape = theta_deg.shape
    r = _rot(theta_deg.flatten(), phi_deg.flatten())
    return r[:, :, 2].reshape(*shape, 3)


def local_to_ecef_dirs(dirs, lat_deg, lon_deg):
    """wgs_84.py:189-220: local (+x east, +y north, +z up) -> ECEF, incl. the 180 deg z flip."""
    r = _rot(90 - lat_deg, 90 - lon_deg).to(dirs.dtype)
    flipped = dirs * torch.tensor([-1.0, -1.0, 1.0], dtype=dirs.dtype)
    return (r @ flipped[..., None])[..., 0]


def build_rays(lat, lon, alt, thetav, phiv, origin_height, tol=10.0, max_iters=20):
    """wgs_84.py:223-290 (get_rays).  Inputs (P, A) float32.  The ray enters the shell at
    altitude `origin_height` and ends on the surface; the entry distance is found by the
    fixed-point update len *= H / alt(len)."""
    x, y, z = geodetic_to_ecef(lat.double(), lon.double(), alt.double())
    surf = torch.stack([x, y, z], dim=-1).float()
    local = view_dirs_local(thetav.double(), phiv.double())
    d = local_to_ecef_dirs(local.view(-1, 3), lat.flatten(), lon.flatten())
    d = -d.view(local.shape)

    lens = (origin_height - alt) / torch.cos(thetav * torch.pi / 180).view(d.shape[:-1]).double()

    def _alt_at(ln):
        p = surf - ln[..., None] * d
        return ecef_to_geodetic(p[..., 0], p[..., 1], p[..., 2])[2]

    alt_now = _alt_at(lens)
    it = 0
    while it < max_iters and (torch.abs(origin_height - alt_now) > tol).any():
        lens = lens * origin_height / alt_now
        alt_now = _alt_at(lens)
        it += 1
    lens = lens.float()
    origins = (surf - d * lens[..., None]).view(-1, 3)
    return origins.float(), d.view(-1, 3).float(), lens.float().flatten()


def valid_ray_mask(origin, direction, rad):
    """wgs_84.py:293-313 (filter_rays)."""
    return ~(origin.isnan().any(dim=1) | direction.isnan().any(dim=1) | rad.isnan())


def normalize_rays(origin, direction, length):
    """wgs_84.py:316-339: bounding box of origins and end points -> (scale, offset)."""
    pts = torch.cat([origin, origin + direction * length[:, None]], dim=0)
    hi = pts.max(dim=0)[0].double()
    lo = pts.min(dim=0)[0].double()
    scale = ((hi - lo).max() / 2).item()
    offset = (hi + lo) / 2
    return torch.clamp((origin - offset) / scale, -1, 1).float(), scale, offset


@dataclass
class HorizontalFrame:
    """The constants captured by the preprocess_coords closure, harp2.py:351-371.

    lat_min/lat_range/lon_min/lon_range are float32 values (they come from float32 granule
    arrays) held here as Python floats, i.e. exactly up-cast to double."""

    scale: float
    offset: tuple  # 3 doubles
    lat_min: float
    lat_range: float
    lon_min: float
    lon_range: float
    shift_lon: bool
    origin_height: float

    @staticmethod
    def from_latlon(lat, lon, scale, offset, origin_height):
        la = lat[~lat.isnan()].float()
        lo = lon[~lon.isnan()].float()
        la_min, la_max = la.min(), la.max()
        lo_min, lo_max = lo.min(), lo.max()
        la_rng, lo_rng = la_max - la_min, lo_max - lo_min
        shift = bool(lo_max > 179 and lo_min < -179)
        if shift:
            lo = lo % 360 - 180
            lo_min, lo_max = lo.min(), lo.max()
            lo_rng = lo_max - lo_min
        return HorizontalFrame(
            float(scale), tuple(float(v) for v in offset), float(la_min), float(la_rng),
            float(lo_min), float(lo_rng), shift, float(origin_height),
        )


def preprocess_horizontal(p, fr: HorizontalFrame):
    """harp2.py:372-386 (preprocess_coords).  `p` is (..., 3) float32 (training) or float64
    (extract).  The multiply by `scale` happens in p's dtype (tensor * python float keeps
    the tensor dtype); the add of the float64 offset promotes to float64."""
    dt = p.dtype
    xyz = p * fr.scale + torch.tensor(fr.offset, dtype=torch.float64)
    lat, lon, alt = ecef_to_geodetic(xyz[..., 0], xyz[..., 1], xyz[..., 2])
    if fr.shift_lon:
        lon = lon % 360 - 180
    lat = 2 * (lat - fr.lat_min) / fr.lat_range - 1
    lon = 2 * (lon - fr.lon_min) / fr.lon_range - 1
    alt = 2 * alt / fr.origin_height - 1
    out = torch.stack([lat, lon, alt], dim=-1).to(dt)
    return torch.clip(out, min=-1, max=1)
