"""Voxel-grid extract dataset (reference: src/atmonr/datasets/harp2_extract.py:189-596).

Only what feeds `pipeline.extract` on the hot path is built here: the point table `xyz`
(float64, (rows*cols*levels, 3), WGS-84 Cartesian) and `idx` (int32), `__getbatch__`, and a
`dump` that stores the extinction grid. Differences from the reference, all outside the hot
path: by default the horizontal layout is a regular lat/lon grid whose spacing
equals `horizontal_step` metres at the scene centre; `layout="vincenty"` (or
ATMONR_EXTRACT_LAYOUT=vincenty) lays the columns out like the reference, along Vincenty geodesics
between the granule's corners (harp2_extract.py:219-330; pinned to the reference class in
tests/test_reference_interchange.py). The reference's DEM lookup only fills the `height` variable of
the output file (the query points are placed above the ELLIPSOID either way); the DEM file is not
shipped, so `height` is zero here. Output is netCDF when `netCDF4` is installed, else `.npz` with the
same variable names.
"""

from __future__ import annotations

import math
from pathlib import Path

import numpy as np
import torch

import os

from atmonr.geospatial.wgs_84 import (WGS_84_A, horizontal_to_cartesian, vincenty_distance,
                                      vincenty_point_along_geodesic)


class HARP2VoxelGridExtractDataset:
    def __init__(self, dataset, horizontal_step: float, alt_step: float, min_alt: float | None = None,
                 max_alt: float | None = None, *args, layout: str | None = None, **kwargs) -> None:
        self.dataset = dataset
        self.layout = layout or os.environ.get("ATMONR_EXTRACT_LAYOUT", "regular")
        if self.layout not in ("regular", "vincenty"):
            raise ValueError(f"unknown extract layout {self.layout!r}")
        self.device = dataset.lat.device
        self.horizontal_step, self.alt_step = float(horizontal_step), float(alt_step)
        self.min_alt = 0.0 if min_alt is None else float(min_alt)
        self.max_alt = float(dataset.config["ray_origin_height"] if max_alt is None else max_alt)
        self.sample_alt = torch.arange(self.min_alt, self.max_alt + self.alt_step / 2, self.alt_step, device=self.device)

        if self.layout == "vincenty":
            self._init_vincenty_grid()
            return
        lat = dataset.lat[~dataset.lat.isnan()]
        lon = dataset.lon[~dataset.lon.isnan()]
        lat_lo, lat_hi, lon_lo, lon_hi = lat.min().item(), lat.max().item(), lon.min().item(), lon.max().item()
        lat_c = 0.5 * (lat_lo + lat_hi)
        dlat = math.degrees(self.horizontal_step / WGS_84_A)
        dlon = dlat / max(math.cos(math.radians(lat_c)), 1e-6)
        rows = max(int((lat_hi - lat_lo) // dlat), 1)
        cols = max(int((lon_hi - lon_lo) // dlon), 1)
        row_lat = lat_hi - (torch.arange(rows, device=self.device, dtype=torch.float64) + 0.5) * dlat  # north first
        col_lon = lon_lo + (torch.arange(cols, device=self.device, dtype=torch.float64) + 0.5) * dlon
        n_alt = self.sample_alt.shape[0]
        self.lat = row_lat[:, None, None].expand(rows, cols, n_alt).contiguous()
        self.lon = col_lon[None, :, None].expand(rows, cols, n_alt).contiguous()
        self.height = torch.zeros((rows, cols), device=self.device, dtype=torch.float64)
        alt = self.sample_alt.double()[None, None].expand(rows, cols, n_alt).contiguous()
        x, y, z = horizontal_to_cartesian(self.lat, self.lon, alt)
        self.shp = self.lat.shape
        self.xyz = torch.stack([x, y, z], dim=-1).view(-1, 3)
        self.idx = torch.arange(self.xyz.shape[0], dtype=torch.int32)

    def _init_vincenty_grid(self) -> None:
        """Columns spaced `horizontal_step` metres apart along geodesics (harp2_extract.py:219-330):
        the four image corners (most extreme valid view of each corner pixel) define a top and a bottom
        edge; evenly spaced stations along both edges are joined by geodesics ("columns"), and every
        column is sampled at evenly spaced fractions of its own length. The grid is centred: what does
        not fit a whole number of steps is split between the two ends."""
        ds = self.dataset
        n_view = ds.view_idx.shape[0]
        lat_img = ds.lat.view(*ds.img_shp, n_view)
        lon_img = ds.lon.view(*ds.img_shp, n_view)
        if not torch.nanmean(lat_img[-1, 0] - lat_img[0, 0]) < 0:
            raise AssertionError("the image must have north at the top")
        east = torch.nanmean(lon_img[0, -1] - lon_img[0, 0]) % 360
        if not (0 < east < 180):
            raise AssertionError("the image must have east on the right")
        lon_ref = torch.nanmean(lon_img)                       # longitudes relative to the scene: no dateline jump
        lon_rel = lon_img - lon_ref
        valid = lambda t: t[~t.isnan()]
        wrap = lambda lon: (lon + 180) % 360 - 180
        corners = {}
        for name, (r, c), lat_pick, lon_pick in (("tl", (0, 0), torch.max, torch.min), ("bl", (-1, 0), torch.min, torch.min),
                                                 ("tr", (0, -1), torch.max, torch.max), ("br", (-1, -1), torch.min, torch.max)):
            if lat_img[r, c].isnan().all() or lon_img[r, c].isnan().all():
                raise AssertionError("a corner pixel of the image has no valid view")
            corners[name] = (lat_pick(valid(lat_img[r, c])), wrap(lon_pick(valid(lon_rel[r, c])) + lon_ref))

        def midpoint(a, b):
            dist, azimuth, _ = vincenty_distance(a, b)
            return vincenty_point_along_geodesic(a, azimuth, dist / 2)[0]

        across, _, _ = vincenty_distance(midpoint(corners["tl"], corners["bl"]), midpoint(corners["tr"], corners["br"]))
        down, _, _ = vincenty_distance(midpoint(corners["tl"], corners["tr"]), midpoint(corners["bl"], corners["br"]))
        rows, cols = int(down // self.horizontal_step), int(across // self.horizontal_step)
        slack_down, slack_across = down % self.horizontal_step, across % self.horizontal_step
        frac_down = (torch.linspace(0, down - slack_down, rows).to(self.device) + slack_down / 2) / down
        frac_across = (torch.linspace(0, across - slack_across, cols).to(self.device) + slack_across / 2) / across
        top_len, top_az, _ = vincenty_distance(corners["tl"], corners["tr"])
        bot_len, bot_az, _ = vincenty_distance(corners["bl"], corners["br"])
        as_vec = lambda v: torch.tensor([float(v)], dtype=torch.float32, device=self.device)
        top, _ = vincenty_point_along_geodesic(torch.stack(corners["tl"]), as_vec(top_az), frac_across * top_len)
        bot, _ = vincenty_point_along_geodesic(torch.stack(corners["bl"]), as_vec(bot_az), frac_across * bot_len)
        col_len, col_az, _ = vincenty_distance(top, bot)
        (grid_lat, grid_lon), _ = vincenty_point_along_geodesic(top[:, None], col_az[None], frac_down[:, None] * col_len[None])
        n_alt = self.sample_alt.shape[0]
        self.height = torch.zeros((rows, cols), device=self.device, dtype=torch.float64)   # no DEM file here
        self.lat = grid_lat[:, :, None].repeat(1, 1, n_alt)
        self.lon = grid_lon[:, :, None].repeat(1, 1, n_alt)
        alt = self.sample_alt[None, None].repeat(rows, cols, 1)
        x, y, z = horizontal_to_cartesian(self.lat.double(), self.lon.double(), alt.double())
        self.shp = self.lat.shape
        self.xyz = torch.stack([x, y, z], dim=-1).view(-1, 3)
        self.idx = torch.arange(self.xyz.shape[0], dtype=torch.int32)

    def __getitem__(self, idx) -> dict[str, torch.Tensor]:
        return {"xyz": self.xyz[idx], "idx": self.idx[idx]}

    def __getbatch__(self, idx: torch.Tensor) -> dict[str, torch.Tensor]:
        return self[idx]

    def __len__(self) -> int:
        return self.xyz.shape[0]

    def dump(self, path: Path, sigma: torch.Tensor) -> None:
        """Store the extinction grid (variable names follow harp2_extract.py:429-596)."""
        rows, cols, n_alt = self.shp
        ext = sigma.detach().float().cpu().numpy().reshape(rows, cols, n_alt, -1)
        xyz = self.xyz.cpu().numpy().reshape(rows, cols, n_alt, 3)
        fields = {
            "extinction_coefficient": ext, "latitude": self.lat[..., 0].cpu().numpy(),
            "longitude": self.lon[..., 0].cpu().numpy(), "height": self.height.cpu().numpy(),
            "altitude": self.sample_alt.cpu().numpy(),
            "x_wgs84": xyz[..., 0], "y_wgs84": xyz[..., 1], "z_wgs84": xyz[..., 2],
        }
        path = Path(path)
        try:
            import netCDF4  # noqa: PLC0415
        except ImportError:
            np.savez_compressed(path.with_suffix(".npz"), **fields)
            return
        with netCDF4.Dataset(path, "w") as nc:  # pragma: no cover - module absent here
            # dimension names of the reference's writer (harp2_extract.py:453-456)
            along, across, vert, bands = "bins_along_track", "bins_across_track", "bins_vertical", "number_of_bands"
            for name, size in ((along, rows), (across, cols), (vert, n_alt), (bands, ext.shape[-1])):
                nc.createDimension(name, size)
            dims = {4: (along, across, vert, bands), 3: (along, across, vert), 2: (along, across), 1: (vert,)}
            for name, arr in fields.items():
                nc.createVariable(name, arr.dtype, dims[arr.ndim])[:] = arr
