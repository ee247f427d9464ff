"""WGS-84 conversions and ray construction (reference: src/atmonr/geospatial/wgs_84.py).

These run ONCE per run at dataset construction (a1, a2), on whatever device the
granule arrays live on, as plain torch expressions in float64 where the reference uses
float64. The per-step, per-sample use of the ECEF -> geodetic conversion (the point
preprocessor) does NOT go through this module: it is fused into the sm_100a sampler kernel
(csrc/device_math.cuh: ecef_to_geodetic).
"""

from __future__ import annotations

import math
import os

import torch

WGS_84_A = 6378137.0
WGS_84_B = 6356752.314245
WGS_84_E = (WGS_84_A**2 - WGS_84_B**2) / (WGS_84_A**2)   # first eccentricity, squared
WGS_84_E2 = (WGS_84_A**2 - WGS_84_B**2) / (WGS_84_B**2)  # second eccentricity, squared
WGS_84_F = (WGS_84_A - WGS_84_B) / WGS_84_A


def horizontal_to_cartesian(lat, lon, alt):
    """(lat, lon in degrees, ellipsoidal height in m) -> ECEF x, y, z. wgs_84.py:24-53."""
    if not (lat.shape == lon.shape == alt.shape):
        raise ValueError("lat, lon, alt must share a shape")
    shape = lat.shape
    phi = lat.flatten() * math.pi / 180
    lam = lon.flatten() * math.pi / 180
    h = alt.flatten()
    sin_phi, cos_phi = torch.sin(phi), torch.cos(phi)
    prime_vertical = WGS_84_A / torch.sqrt(1 - (WGS_84_E * sin_phi**2))
    x = (prime_vertical + h) * cos_phi * torch.cos(lam)
    y = (prime_vertical + h) * cos_phi * torch.sin(lam)
    z = (prime_vertical * (1 - WGS_84_E) + h) * sin_phi
    return x.view(shape), y.view(shape), z.view(shape)


def cartesian_to_horizontal(x, y, z):
    """ECEF -> (lat deg, lon deg, height m) with ONE Bowring iteration; the height is recovered
    as x / (cos lat cos lon) - N. wgs_84.py:56-97 (kept literal: same accuracy, same
    singularities)."""
    if not (x.shape == y.shape == z.shape):
        raise ValueError("x, y, z must share a shape")
    shape = x.shape
    x, y, z = x.flatten(), y.flatten(), z.flatten()
    lam = torch.atan2(y, x)
    horiz = torch.sqrt(x**2 + y**2)
    param_lat = torch.atan2(z / horiz, torch.zeros_like(x) + WGS_84_A / WGS_84_B)
    phi = torch.atan2(
        z + (WGS_84_E2 * WGS_84_B) * (torch.sin(param_lat) ** 3),
        horiz - (WGS_84_E * WGS_84_A) * (torch.cos(param_lat) ** 3),
    )
    prime_vertical = WGS_84_A / torch.sqrt(1 - (WGS_84_E * torch.sin(phi) ** 2))
    height = x / (torch.cos(phi) * torch.cos(lam)) - prime_vertical
    to_deg = lambda t: t.view(shape) * 180 / torch.pi
    return to_deg(phi), to_deg(lam), height.view(shape)


def horizontal_coords_to_rot_mtx(theta, phi):
    """Rotation matrices (n,3,3) from zenith/azimuth in degrees; both angles are negated to match
    the 3-D rotation convention. wgs_84.py:100-132."""
    if theta.dim() != 1 or theta.shape != phi.shape:
        raise ValueError("theta and phi must be 1-D and the same length")
    t = -theta * torch.pi / 180
    p = -phi * torch.pi / 180
    st, ct, sp, cp = torch.sin(t), torch.cos(t), torch.sin(p), torch.cos(p)
    zero = torch.zeros_like(t)
    row0 = torch.stack([cp, -sp * ct, sp * st], dim=1)
    row1 = torch.stack([sp, cp * ct, -cp * st], dim=1)
    row2 = torch.stack([zero, st, ct], dim=1)
    return torch.stack([row0, row1, row2], dim=1)


def horizontal_coords_to_dirvecs(theta, phi):
    """Unit vectors for (zenith, azimuth) in the local +z-up frame. wgs_84.py:135-160."""
    if theta.shape != phi.shape:
        raise ValueError("theta and phi must share a shape")
    shape = tuple(theta.shape)
    rot = horizontal_coords_to_rot_mtx(theta.flatten(), phi.flatten())
    up = torch.zeros((rot.shape[0], 3, 1), dtype=rot.dtype, device=rot.device)
    up[:, 2] = 1
    return (rot @ up).view(*shape, 3)


def dirvecs_to_horizontal_coords(dirs):
    """Inverse of horizontal_coords_to_dirvecs. wgs_84.py:163-186."""
    d = dirs.view(-1, 3)
    theta = torch.atan2(torch.linalg.norm(d[..., :2]), d[..., 2])
    phi = -torch.atan2(d[..., 0], -d[..., 1])
    return (theta * 180 / torch.pi) % 360, (phi * 180 / torch.pi) % 360 - 180


def compose_dirs_and_surface_normals(dirs, lat, lon):
    """Local-frame directions -> ECEF directions. wgs_84.py:189-220 (includes the 180 degree
    turn about z between the scene convention and the WGS convention)."""
    rot = horizontal_coords_to_rot_mtx(90 - lat, 90 - lon).to(dtype=dirs.dtype)
    turn = torch.tensor([-1.0, -1.0, 1.0], dtype=dirs.dtype, device=dirs.device)
    return (rot @ (dirs * turn)[..., None])[..., 0]


def get_rays(lat, lon, alt, thetav, phiv, ray_origin_height, tol: float = 10.0, max_iters: int = 20):
    """Ray entry (top of the shell at `ray_origin_height`) and exit (surface) for every
    pixel/view. wgs_84.py:223-290. Returns origins (P*A,3), directions (P*A,3), lengths (P*A,).

    CUDA inputs are computed by the library's ray-setup kernels (`atmonr_get_rays`, csrc/rays.cu: same
    dtype flow, same chunk-wide refinement loop); ATMONR_NATIVE_RAYS=0 keeps the torch expressions
    below on the device as well (the cross-check of tests/test_zz_gpu_rays.py). CPU inputs (dataset
    interchange tests in the build container) always take the torch expressions."""
    if lat.is_cuda and os.environ.get("ATMONR_NATIVE_RAYS", "1") != "0":
        from atmonr.native import ops
        return ops.get_rays(lat, lon, alt, thetav, phiv, ray_origin_height, tol, max_iters)
    x, y, z = horizontal_to_cartesian(lat.double(), lon.double(), alt.double())
    surface = torch.stack([x, y, z], dim=-1).float()
    local = horizontal_coords_to_dirvecs(thetav.double(), phiv.double())
    dirs = -compose_dirs_and_surface_normals(local.view(-1, 3), lat.flatten(), lon.flatten()).view(local.shape)

    lens = (ray_origin_height - alt) / torch.cos(thetav * torch.pi / 180).view(dirs.shape[:-1]).double()

    def height_at(length):
        p = surface - length[..., None] * dirs
        return cartesian_to_horizontal(p[..., 0], p[..., 1], p[..., 2])[2]

    height = height_at(lens)
    n_iter = 0
    while n_iter < max_iters and (torch.abs(ray_origin_height - height) > tol).any():
        lens = lens * ray_origin_height / height
        height = height_at(lens)
        n_iter += 1
    lens = lens.float()
    origins = (surface - dirs * lens[..., None]).view(-1, 3)
    return origins.float(), dirs.view(-1, 3).float(), lens.float().flatten()


def filter_rays(ray_origin, ray_dir, ray_rad):
    """Mask of rays whose origin, direction and radiance hold no NaN. wgs_84.py:293-313. CUDA inputs:
    `atmonr_filter_rays` (csrc/rays.cu); ATMONR_NATIVE_RAYS=0 / CPU inputs: the torch expressions."""
    if ray_origin.is_cuda and os.environ.get("ATMONR_NATIVE_RAYS", "1") != "0":
        from atmonr.native import ops
        return ops.filter_rays(ray_origin, ray_dir, ray_rad)
    bad = ray_origin.isnan().any(dim=1) | ray_dir.isnan().any(dim=1) | ray_rad.isnan()
    return ~bad


def normalize_rays(ray_origin, ray_dir, ray_len):
    """Scale/offset that maps origins and end points into [-1,1]^3. wgs_84.py:316-339. CUDA inputs:
    `atmonr_ray_extent` + `atmonr_normalize_origins` (csrc/rays.cu), bit for bit the expressions below."""
    if ray_origin.is_cuda and os.environ.get("ATMONR_NATIVE_RAYS", "1") != "0":
        from atmonr.native import ops
        return ops.normalize_rays(ray_origin, ray_dir, ray_len)
    ends = torch.cat([ray_origin, ray_origin + ray_dir * ray_len[:, None]], dim=0)
    hi = ends.max(dim=0)[0].double()
    lo = ends.min(dim=0)[0].double()
    scale = ((hi - lo).max() / 2).item()
    offset = (hi + lo) / 2
    return torch.clamp((ray_origin - offset) / scale, -1, 1).float(), scale, offset


# ---- Vincenty's formulae on the WGS-84 ellipsoid (wgs_84.py:342-575) ------------------------------
# Used by the reference's Vincenty voxel-grid layout (datasets/harp2_extract.py:189-348), which this
# build replaces by a regular grid; the two functions keep the reference's conventions so that layout
# code written against them keeps working: angles in degrees in, distances in metres, the inverse
# problem returns both azimuths in DEGREES, the direct problem returns the destination in degrees
# (longitude not wrapped) and the arrival azimuth in RADIANS, and a stalled iteration raises Warning.
def _reduced(lat_deg):
    """Reduced latitude U: tan U = (1 - f) tan(lat)."""
    return torch.atan((1 - WGS_84_F) * torch.tan(lat_deg * torch.pi / 180))


def _vincenty_series(cos2_alpha):
    """Vincenty's A and B as functions of u^2 = cos^2(alpha) (a^2 - b^2) / b^2."""
    u2 = cos2_alpha * (WGS_84_A**2 - WGS_84_B**2) / WGS_84_B**2
    big_a = 1 + (u2 / 16384) * (4096 + u2 * (-768 + u2 * (320 - 175 * u2)))
    big_b = (u2 / 1024) * (256 + u2 * (-128 + u2 * (74 - 47 * u2)))
    return big_a, big_b


def _delta_sigma(big_b, sin_s, cos_s, cos_2sm):
    inner = cos_s * (-1 + 2 * cos_2sm**2) - (1 / 6) * big_b * cos_2sm * (-3 + 4 * sin_s**2) * (-3 + 4 * cos_2sm**2)
    return big_b * sin_s * (cos_2sm + (1 / 4) * big_b * inner)


def _longitude_term(sin_alpha, cos2_alpha, sigma, sin_s, cos_s, cos_2sm):
    """(1 - C) f sin(alpha) [sigma + C sin(sigma) (cos 2sigma_m + C cos(sigma) (-1 + 2 cos^2 2sigma_m))]"""
    c = (WGS_84_F / 16) * cos2_alpha * (4 + WGS_84_F * (4 - 3 * cos2_alpha))
    return (1 - c) * WGS_84_F * sin_alpha * (sigma + c * sin_s * (cos_2sm + c * cos_s * (-1 + 2 * cos_2sm**2)))


def vincenty_distance(latlon1, latlon2, tol: float = 1e-12, max_iters: int = 10):
    """Inverse problem: geodesic distance (m) and the forward azimuths (degrees) at both ends.
    latlon1 / latlon2: (lat, lon) tuples or (2, ...) tensors, degrees. wgs_84.py:342-449."""
    if not isinstance(tol, float) or not isinstance(max_iters, int):
        raise AssertionError("tol must be a float and max_iters an int")
    u1, u2 = _reduced(latlon1[0]), _reduced(latlon2[0])
    su1, cu1, su2, cu2 = torch.sin(u1), torch.cos(u1), torch.sin(u2), torch.cos(u2)
    dlon = latlon2[1] * torch.pi / 180 - latlon1[1] * torch.pi / 180
    lam, n_iter = dlon, 0
    while True:
        if n_iter > max_iters:
            raise Warning(f"Exceeded {max_iters} iterations without lambda changing by less than {tol:.1e}")
        sl, cl = torch.sin(lam), torch.cos(lam)
        sin_s = torch.sqrt((cu2 * sl) ** 2 + (cu1 * su2 - su1 * cu2 * cl) ** 2)
        cos_s = su1 * su2 + cu1 * cu2 * cl
        sigma = torch.atan2(sin_s, cos_s)
        sin_alpha = cu1 * cu2 * sl / sin_s
        cos2_alpha = 1 - sin_alpha**2
        cos_2sm = cos_s - (2 * su1 * su2) / cos2_alpha
        new = dlon + _longitude_term(sin_alpha, cos2_alpha, sigma, sin_s, cos_s, cos_2sm)
        step, lam, n_iter = new - lam, new, n_iter + 1
        if not bool((torch.abs(step) > tol).any()):
            break
    big_a, big_b = _vincenty_series(cos2_alpha)
    s = WGS_84_B * big_a * (sigma - _delta_sigma(big_b, sin_s, cos_s, cos_2sm))
    sl, cl = torch.sin(lam), torch.cos(lam)
    alpha1 = torch.atan2(cu2 * sl, cu1 * su2 - su1 * cu2 * cl)
    alpha2 = torch.atan2(cu1 * sl, -su1 * cu2 + cu1 * su2 * cl)
    return s, alpha1 * 180 / math.pi, alpha2 * 180 / math.pi


def vincenty_point_along_geodesic(latlon1, alpha1, s, tol: float = 1e-6, max_iters: int = 10):
    """Direct problem: the point `s` metres along the geodesic leaving latlon1 (degrees) at azimuth
    alpha1 (degrees). Returns ((lat2, lon2) in degrees, same container kind as latlon1; arrival azimuth in
    radians). wgs_84.py:452-575."""
    if not isinstance(alpha1, torch.Tensor) or not isinstance(s, torch.Tensor):
        raise AssertionError("alpha1 and s must be tensors")
    if not isinstance(tol, float) or not isinstance(max_iters, int):
        raise AssertionError("tol must be a float and max_iters an int")
    lon1 = latlon1[1] * torch.pi / 180
    az = alpha1 * torch.pi / 180
    u1 = _reduced(latlon1[0])
    su1, cu1, saz, caz = torch.sin(u1), torch.cos(u1), torch.sin(az), torch.cos(az)
    sigma1 = torch.atan2(torch.tan(u1), caz)
    sin_alpha = cu1 * saz
    cos2_alpha = 1 - sin_alpha**2
    big_a, big_b = _vincenty_series(cos2_alpha)
    first = s / (WGS_84_B * big_a)
    sigma, n_iter = first, 0
    while True:
        if n_iter > max_iters:
            raise Warning(f"Exceeded {max_iters} iterations without sigma changing by less than {tol:.1e}")
        cos_2sm = torch.cos(2 * sigma1 + sigma)
        new = first + _delta_sigma(big_b, torch.sin(sigma), torch.cos(sigma), cos_2sm)
        step, sigma, n_iter = new - sigma, new, n_iter + 1
        if not bool((torch.abs(step) > tol).any()):
            break
    ss, cs = torch.sin(sigma), torch.cos(sigma)
    lat2 = torch.atan2(su1 * cs + cu1 * ss * caz,
                       (1 - WGS_84_F) * torch.sqrt(sin_alpha**2 + (su1 * ss - cu1 * cs * caz) ** 2))
    lam = torch.atan2(ss * saz, cu1 * cs - su1 * ss * caz)
    lon2 = lam - _longitude_term(sin_alpha, cos2_alpha, sigma, ss, cs, cos_2sm) + lon1
    alpha2 = torch.atan2(sin_alpha, -su1 * ss + cu1 * cs * caz)
    lat2, lon2 = lat2 * 180 / torch.pi, lon2 * 180 / torch.pi
    return ((lat2, lon2) if isinstance(latlon1, tuple) else torch.stack([lat2, lon2])), alpha2
