"""Voxel-grid extract dataset (reference: src/atmonr/datasets/harp2_extract.py:189-596).

Only what feeds `pipeline.extract` on the hot path is built here: the point table `xyz`
(float64, (rows*cols*levels, 3), WGS-84 Cartesian) and `idx` (int32), `__getbatch__`, and a
`dump` that stores the extinction grid. Differences from the reference, all outside the hot
path: the horizontal layout is a regular lat/lon grid whose spacing equals
`horizontal_step` metres at the scene centre (the reference spaces columns along Vincenty
geodesics and lifts them by a DEM file that is not shipped); output is netCDF when `netCDF4` is
installed, else `.npz` with the same variable names.
"""

from __future__ import annotations

import math
from pathlib import Path

import numpy as np
import torch

from atmonr.geospatial.wgs_84 import WGS_84_A, horizontal_to_cartesian


class HARP2VoxelGridExtractDataset:
    def __init__(self, dataset, horizontal_step: float, alt_step: float, min_alt: float | None = None,
                 max_alt: float | None = None, *args, **kwargs) -> None:
        self.dataset = dataset
        self.device = dataset.lat.device
        self.horizontal_step, self.alt_step = float(horizontal_step), float(alt_step)
        self.min_alt = 0.0 if min_alt is None else float(min_alt)
        self.max_alt = float(dataset.config["ray_origin_height"] if max_alt is None else max_alt)
        self.sample_alt = torch.arange(self.min_alt, self.max_alt + self.alt_step / 2, self.alt_step, device=self.device)

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
            for name, size in (("rows", rows), ("cols", cols), ("levels", n_alt), ("bands", ext.shape[-1])):
                nc.createDimension(name, size)
            dims = {4: ("rows", "cols", "levels", "bands"), 3: ("rows", "cols", "levels"), 2: ("rows", "cols"), 1: ("levels",)}
            for name, arr in fields.items():
                nc.createVariable(name, arr.dtype, dims[arr.ndim])[:] = arr
