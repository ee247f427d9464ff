"""More golden vectors from the REFERENCE's own code (same rules as make_golden.py: run in the build
container only, outputs committed): the two sampler-side functions that are off the default configs
(SURVEY 8f-4), `sample_biased_bins` and `append_heights`.

    python tests/golden/make_golden_extra.py   ->   tests/golden/reference_vectors_extra.npz
"""

import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import import_reference, synthetic_geometry  # noqa: E402


def main():
    import_reference()
    from atmonr import samplers
    from atmonr.geospatial import wgs_84

    out = {}
    lat, lon, alt, thetav, phiv = synthetic_geometry()
    origins, dirs, lens = wgs_84.get_rays(lat, lon, alt, thetav, phiv, 20000.0)
    origins_n, scale, offset = wgs_84.normalize_rays(origins, dirs, lens)
    batch = {"origin": origins_n[:40], "dir": dirs[:40], "len": lens[:40] / scale}
    for tag, alpha in (("a0", 0.0), ("a35", 0.35), ("a9", 0.9)):
        torch.manual_seed(11)
        pts, z = samplers.sample_biased_bins(batch, 24, 20000.0, alpha)
        out[f"bias_{tag}_pts"], out[f"bias_{tag}_z"] = pts.numpy(), z.numpy()
    torch.manual_seed(11)
    out["bias_u"] = torch.rand((40, 24)).numpy()          # the draws the calls above consumed
    for k, v in batch.items():
        out["bias_" + k] = v.numpy()
    torch.manual_seed(12)
    pts, _ = samplers.sample_uniform_bins(batch, n_bins=16)
    out["ah_pts"] = pts.numpy()
    out["ah_out"] = samplers.append_heights(pts, 20000.0, scale, offset).numpy()
    out["ah_scale"], out["ah_offset"] = np.float64(scale), offset.numpy()
    # graphics_utils.voxel_traversal: random segments in 2-D and 3-D, axis-aligned ones, and segments
    # that start and end in one voxel; compared as SETS of visited voxels (sorted unique rows)
    from atmonr import graphics_utils
    g = torch.Generator().manual_seed(21)
    for tag, dim in (("2d", 2), ("3d", 3)):
        a = torch.rand(60, dim, generator=g) * 20 - 4
        b = torch.rand(60, dim, generator=g) * 20 - 4
        b[:6, 0] = a[:6, 0]                       # no motion along x
        b[6:10] = a[6:10] + 0.01 * (torch.rand(4, dim, generator=g) - 0.5)   # (almost always) one voxel
        reg = graphics_utils.voxel_traversal(a.clone(), b.clone(), unique_only=True)
        out[f"vox_{tag}_u"], out[f"vox_{tag}_end"] = a.numpy(), b.numpy()
        out[f"vox_{tag}_set"] = torch.unique(reg, dim=0).numpy()
        full = graphics_utils.voxel_traversal(a.clone(), b.clone(), unique_only=False)
        out[f"vox_{tag}_visits"] = np.int64(full.shape[0])
    # geospatial/spherical.py: the spherical-Earth helpers of the global-grid layout
    from atmonr.geospatial import spherical
    xyz = origins[:50].clone()
    out["sph_in"] = xyz.numpy()
    out["sph_fwd"] = spherical.wgs_84_to_spherical(xyz.clone()).numpy()
    out["sph_back"] = spherical.spherical_to_wgs84(spherical.wgs_84_to_spherical(xyz.clone())).numpy()
    out["sph_stretch"] = spherical.stretch_above_sea_level(spherical.wgs_84_to_spherical(xyz.clone()), 12.0).numpy()
    # Vincenty inverse / direct problems (wgs_84.py:342-575), float64 inputs
    g = torch.Generator().manual_seed(33)
    la1 = torch.rand(40, generator=g, dtype=torch.float64) * 120 - 60
    lo1 = torch.rand(40, generator=g, dtype=torch.float64) * 300 - 150
    la2 = la1 + torch.rand(40, generator=g, dtype=torch.float64) * 8 - 4
    lo2 = lo1 + torch.rand(40, generator=g, dtype=torch.float64) * 8 - 4
    s_, a1, a2 = wgs_84.vincenty_distance((la1, lo1), (la2, lo2))
    out["vin_ll1"], out["vin_ll2"] = torch.stack([la1, lo1]).numpy(), torch.stack([la2, lo2]).numpy()
    out["vin_s"], out["vin_a1"], out["vin_a2"] = s_.numpy(), a1.numpy(), a2.numpy()
    (la3, lo3), a3 = wgs_84.vincenty_point_along_geodesic((la1, lo1), a1, s_ * 0.37)
    out["vin_d_lat"], out["vin_d_lon"], out["vin_d_a2"] = la3.numpy(), lo3.numpy(), a3.numpy()
    ll4, _ = wgs_84.vincenty_point_along_geodesic(torch.stack([la1, lo1]), a1, s_)
    out["vin_d_full"] = ll4.numpy()
    np.savez_compressed(HERE / "reference_vectors_extra.npz", **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
