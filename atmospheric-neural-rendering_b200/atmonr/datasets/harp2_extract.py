"""Voxel-grid extract dataset (reference: src/atmonr/datasets/harp2_extract.py:189-596).

Only what feeds `pipeline.extract` on the hot path is built here: the point table `xyz`
(float64, (rows*cols*levels, 3), WGS-84 Cartesian) and `idx` (int32), `__getbatch__`, and a
`dump` that stores the extinction grid. The columns are laid out like the reference's, along Vincenty
geodesics between the granule's corners (harp2_extract.py:219-330; pinned to the reference class in
tests/test_reference_interchange.py); `layout="regular"` (or ATMONR_EXTRACT_LAYOUT=regular) gives a
regular lat/lon grid whose spacing equals `horizontal_step` metres at the scene centre instead (the dense
grid SURVEY 8d specifies for the extraction benchmark; it also serves scenes whose corner pixels have no
valid view, which the reference's layout rejects). Differences from the reference, outside the hot path
(SURVEY 8f-1/3): the reference's DEM lookup only fills the `height` variable of the output file (the query
points are placed above the ELLIPSOID either way); the DEM file is not shipped, so `height` is zero here.
Output is netCDF when `netCDF4` is installed, else `.npz` with the same variable names.

The other three coordinate modes of scripts/extract.py are point-table LAYOUTS around the
same `pipeline.extract` call: `HARP2L1CExtractDataset` (harp2_extract.py:115-186: the 5 km L1C bin
grid), `HARP2EarthCAREExtractDataset` (:599-791: the curtain under an EarthCARE ATLID track) and
`HARP2GlobalGridExtractDataset` (:794-946: the voxels of a global spherical-Earth grid that the
granule's rays cross). Their auxiliary inputs are read with netCDF4 / h5py when those modules are
installed, from an `.npz` file with the same variable paths as keys otherwise, and for a `synthetic:`
granule from the generator's stand-ins (datasets/granule.py). All three are pinned to the reference's
classes in tests/test_reference_interchange.py.
"""

from __future__ import annotations

import math
from pathlib import Path

import numpy as np
import torch

import os

from atmonr.geospatial.spherical import spherical_to_wgs84, stretch_above_sea_level, wgs_84_to_spherical
from atmonr.geospatial.wgs_84 import (WGS_84_A, cartesian_to_horizontal, horizontal_to_cartesian,
                                      vincenty_distance, vincenty_point_along_geodesic)

_CHUNK_SIZE = int(3e4)      # rays per voxel-traversal call (harp2_extract.py:34)


def _read_aux(path: Path, names: list[str]) -> dict[str, np.ndarray]:
    """Variables `names` ("group/variable" paths) of an auxiliary product as numpy arrays, invalid
    values as NaN: an `.npz` file with those paths as keys, else netCDF (`.nc`, netCDF4) / HDF5 (`.h5`,
    h5py) when the module is installed."""
    path = Path(path)
    npz = path if path.suffix == ".npz" else path.with_suffix(".npz")
    if npz.exists():
        with np.load(npz, allow_pickle=False) as f:
            return {n: np.asarray(f[n]) for n in names}
    if not path.exists():
        raise FileNotFoundError(f"{path} (or {npz.name} with the keys {names}) not found; downloads are outside this build")
    if path.suffix == ".nc":
        try:
            import netCDF4  # noqa: PLC0415
        except ImportError as err:
            raise ImportError(f"reading {path} needs the `netCDF4` module (or an .npz copy)") from err
        with netCDF4.Dataset(path) as nc:  # pragma: no cover - module absent here
            return {n: nc[n][:].filled(fill_value=np.nan) for n in names}
    try:
        import h5py  # noqa: PLC0415
    except ImportError as err:
        raise ImportError(f"reading {path} needs the `h5py` module (or an .npz copy)") from err
    with h5py.File(path) as f:  # pragma: no cover - module absent here
        return {n: f[n][()] for n in names}


def _save_fields(path: Path, fields: dict[str, np.ndarray], dims: dict[str, tuple[str, ...]],
                 attrs: dict | None = None) -> None:
    """netCDF with the reference's dimension / variable names when `netCDF4` is installed, else `.npz`."""
    path = Path(path)
    try:
        import netCDF4  # noqa: PLC0415
    except ImportError:
        np.savez_compressed(path.with_suffix(".npz"), **fields, **{f"attr_{k}": np.asarray(v) for k, v in (attrs or {}).items()})
        return
    with netCDF4.Dataset(path, "w") as nc:  # pragma: no cover - module absent here
        for name, arr in fields.items():
            for d, size in zip(dims[name], arr.shape):
                if d not in nc.dimensions:
                    nc.createDimension(d, size)
            nc.createVariable(name, arr.dtype, dims[name])[:] = arr
        for k, v in (attrs or {}).items():
            setattr(nc, k, v)


class _ExtractTable:
    """harp2_extract.py:38-68: the point table every layout fills (`xyz` float64 WGS-84 Cartesian,
    `idx` int32) and the two accessors the loader uses."""
    xyz: torch.Tensor
    idx: torch.Tensor

    def __getitem__(self, idx) -> dict[str, torch.Tensor]:
        return {"xyz": self.xyz[idx], "idx": self.idx[idx]}

    def __getbatch__(self, idx: torch.Tensor) -> dict[str, torch.Tensor]:
        return self[idx]

    def __len__(self) -> int:
        return self.xyz.shape[0]


class HARP2VoxelGridExtractDataset(_ExtractTable):
    def __init__(self, dataset, horizontal_step: float, alt_step: float, min_alt: float | None = None,
                 max_alt: float | None = None, *args, layout: str | None = None, **kwargs) -> None:
        self.dataset = dataset
        self.layout = layout or os.environ.get("ATMONR_EXTRACT_LAYOUT", "vincenty")
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

    def dump(self, path: Path, sigma: torch.Tensor) -> None:
        """Store the extinction grid (variable names follow harp2_extract.py:429-596)."""
        rows, cols, n_alt = self.lat.shape          # (the L1C layout keeps the 2-D bin shape in `shp`)
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


class HARP2L1CExtractDataset(HARP2VoxelGridExtractDataset):
    """harp2_extract.py:115-186: voxel columns over the bins of the granule's level-1C grid (5 km, map
    projected: evenly spaced, unlike the view-dependent L1B geolocation), at the user's altitude levels
    above the ELLIPSOID (the L1C `height` only goes into the output file). `dump` is the voxel grid's
    (harp2_extract.py:100-112 shares `_extract_to_netCDF` between the two as well)."""

    def __init__(self, dataset, alt_step: float, min_alt: float | None = None, max_alt: float | None = None,
                 *args, l1c_path: Path | None = None, **kwargs) -> None:
        self.dataset, self.layout = dataset, "l1c"
        self.device = dataset.lat.device
        self.alt_step = float(alt_step)
        self.min_alt = 0.0 if min_alt is None else float(min_alt)
        self.max_alt = float(dataset.config["ray_origin_height"] if max_alt is None else max_alt)
        self.sample_alt = torch.arange(self.min_alt, self.max_alt + self.alt_step / 2, self.alt_step, device=self.device)
        names = ["geolocation_data/latitude", "geolocation_data/longitude", "geolocation_data/height"]
        if l1c_path is None and str(dataset.filename).startswith("synthetic"):
            geo = dataset.granule.l1c_geolocation()
            raw = {n: geo[n.split("/")[1]] for n in names}
        else:
            if l1c_path is None:            # harp2_extract.py:141-146: the L1C granule of the same scene
                sensor, timestamp, _, version, _ = str(dataset.filename).split(".")
                l1c_path = Path("data/HARP2_L1C") / f"{sensor}.{timestamp}.L1C.{version}.5km.nc"
            raw = _read_aux(Path(l1c_path), names)
        # fill -> NaN (done by the reader), y-axis flipped so north is the first row, on the dataset's device
        lat, lon, self.height = (torch.from_numpy(np.ascontiguousarray(raw[n][::-1])).to(self.device) for n in names)
        n_alt = self.sample_alt.shape[0]
        self.lat = lat[:, :, None].repeat((1, 1, n_alt))
        self.lon = lon[:, :, None].repeat((1, 1, n_alt))
        alt = self.sample_alt[None, None].repeat(lat.shape[0], lat.shape[1], 1)
        x, y, z = horizontal_to_cartesian(self.lat.double(), self.lon.double(), alt.double())
        self.shp = tuple(lat.shape)                 # the reference keeps the 2-D bin shape here
        self.xyz = torch.stack([x, y, z], dim=-1).view(-1, 3)
        self.idx = torch.arange(self.xyz.shape[0], dtype=torch.int32)


class HARP2EarthCAREExtractDataset(_ExtractTable):
    """harp2_extract.py:599-791: the curtain under an EarthCARE ATLID track (product `ATL_EBD_2A`): one
    point per (profile, joint-standard-grid range bin), optionally only the profiles
    [earthcare_range[0], earthcare_range[1]); range bins are kept when EVERY kept profile has them
    strictly between the ellipsoid and `ray_origin_height`."""

    def __init__(self, dataset, earthcare_filename: str, earthcare_range: list[int] | None = None, *args, **kwargs) -> None:
        if not (earthcare_range is None or (len(earthcare_range) == 2 and earthcare_range[1] > earthcare_range[0])):
            raise AssertionError("earthcare_range must be [start, end) with end > start")
        self.dataset, self.device = dataset, dataset.lat.device
        self.earthcare_filename, self.earthcare_range = earthcare_filename, earthcare_range
        if str(earthcare_filename).startswith("synthetic"):
            track = dataset.granule.earthcare_track()
            file_type, alt, lat, lon = track["file_type"], track["height"], track["latitude"], track["longitude"]
        else:
            names = ["HeaderData/FixedProductHeader/File_Type", "ScienceData/height", "ScienceData/latitude", "ScienceData/longitude"]
            raw = _read_aux(Path("data") / "EarthCARE" / earthcare_filename, names)
            file_type = raw[names[0]]
            file_type = file_type.item() if isinstance(file_type, np.ndarray) else file_type
            file_type = file_type.decode() if isinstance(file_type, bytes) else str(file_type)
            alt, lat, lon = raw[names[1]], raw[names[2]], raw[names[3]]
        if file_type != "ATL_EBD_2A":
            raise NotImplementedError(f"Extraction currently only supports ATL_EBD_2A, not supported for '{file_type}'.")
        self.alt = np.asarray(alt)
        self.lat = np.repeat(np.asarray(lat)[:, None], self.alt.shape[1], axis=1)
        self.lon = np.repeat(np.asarray(lon)[:, None], self.alt.shape[1], axis=1)
        if earthcare_range is not None:
            keep = slice(int(earthcare_range[0]), int(earthcare_range[1]))
            self.lat, self.lon, self.alt = self.lat[keep], self.lon[keep], self.alt[keep]
        in_shell = (self.alt > 0).all(axis=0) * (self.alt < dataset.config["ray_origin_height"]).all(axis=0)
        self.lat, self.lon, self.alt = self.lat[:, in_shell], self.lon[:, in_shell], self.alt[:, in_shell]
        self.shp = self.lat.shape
        on_device = lambda a: torch.from_numpy(a.flatten()).to(self.device)
        self.xyz = torch.stack(horizontal_to_cartesian(on_device(self.lat), on_device(self.lon), on_device(self.alt)), dim=1)
        self.idx = torch.arange(self.xyz.shape[0], dtype=torch.int32)

    def dump(self, path: Path, sigma: torch.Tensor) -> None:
        """Variables and dimensions of harp2_extract.py:676-791. (The reference's writer stores the
        LATITUDE in its `longitude` variable, :746; the longitude is written here.)"""
        n_bands = sigma.shape[-1]
        xyz = self.xyz.view(*self.shp, 3).cpu().numpy().astype(np.float32)
        fields = {"latitude": self.lat[..., 0].astype(np.float64), "longitude": self.lon[..., 0].astype(np.float64),
                  "height": self.alt.astype(np.float64),
                  "extinction_coefficient": sigma.detach().float().cpu().numpy().reshape(*self.shp, n_bands),
                  "x_wgs84": xyz[..., 0], "y_wgs84": xyz[..., 1], "z_wgs84": xyz[..., 2]}
        curtain = ("along_track", "JSG_height")
        dims = {"latitude": curtain[:1], "longitude": curtain[:1], "height": curtain,
                "extinction_coefficient": (*curtain, "number_of_bands"), "x_wgs84": curtain, "y_wgs84": curtain, "z_wgs84": curtain}
        attrs = {"neural_rendering_scene_scale": float(self.dataset.scale), "ray_origin_height": float(self.dataset.config["ray_origin_height"])}
        for k, axis in zip("xyz", range(3)):
            attrs[f"neural_rendering_scene_offset_{k}"] = float(self.dataset.offset[axis])
        if isinstance(self.earthcare_range, list):
            attrs["earthcare_start_idx"], attrs["earthcare_end_idx"] = int(self.earthcare_range[0]), int(self.earthcare_range[1])
        _save_fields(path, fields, dims, attrs)


class HARP2GlobalGridExtractDataset(_ExtractTable):
    """harp2_extract.py:794-946: for large-scale visualisation, the voxels of a GLOBAL grid (spherical
    Earth, `scale` scene units per metre, voxels of `grid_res` scene units) that the granule's rays
    cross, with everything above sea level stretched radially by `vstretch` before the voxels are
    chosen and un-stretched afterwards; in every z-layer of voxels the `lon_crop` fraction of the
    longitude range is dropped at both ends. `xyz` = the voxel centres back in WGS-84 Cartesian metres
    (float32, as the reference computes them), `voxels` = their integer grid indices.

    The reference's last step (:896, `cull = alt <= 0 + alt > H`) is a chained comparison of tensors and
    raises for any grid of more than one voxel; what its comment states is done here: voxel centres at
    or below the ellipsoid, or above `ray_origin_height`, are dropped."""

    def __init__(self, dataset, scale: float, grid_res: float, vstretch: float | None = None, lon_crop: float = 0.05,
                 *args, **kwargs) -> None:
        from atmonr.graphics_utils import voxel_traversal  # noqa: PLC0415
        vstretch = 1 if vstretch is None else vstretch
        if not vstretch >= 1:
            raise AssertionError("vstretch must be >= 1")
        self.dataset, self.device = dataset, dataset.lat.device
        self.scale, self.grid_res, self.vstretch = scale, grid_res, vstretch
        to_grid = self.scale / self.grid_res
        top = wgs_84_to_spherical(dataset.ray_origin)
        bottom = wgs_84_to_spherical(dataset.ray_origin + dataset.ray_dir * dataset.ray_len[:, None])
        top = stretch_above_sea_level(top, self.vstretch) * to_grid
        bottom = stretch_above_sea_level(bottom, self.vstretch) * to_grid
        cells = torch.zeros((0, 3), device=self.device)
        for start in range(0, top.shape[0], _CHUNK_SIZE):
            seen = voxel_traversal(top[start:start + _CHUNK_SIZE], bottom[start:start + _CHUNK_SIZE], unique_only=False)
            cells = torch.unique(torch.cat([cells, seen], dim=0), dim=0, sorted=False)     # per chunk: bounds the memory
        centres = (cells.float() + 0.5) * (self.grid_res / self.scale)
        lon = torch.atan2(centres[..., 1], centres[..., 0])
        kept = []
        for z in torch.unique(centres[..., 2]):
            layer = centres[..., 2] == z
            lon_layer = lon[layer]
            span = lon_layer.max() - lon_layer.min()
            inside = (lon_layer > lon_layer.min() + lon_crop * span) * (lon_layer < lon_layer.max() - lon_crop * span)
            kept.append(centres[layer][inside])
        centres = torch.cat(kept, dim=0) if kept else centres
        self.voxels = (centres * to_grid).to(dtype=torch.int32)
        xyz = spherical_to_wgs84(stretch_above_sea_level(centres, 1 / self.vstretch))
        _, _, alt = cartesian_to_horizontal(xyz[..., 0], xyz[..., 1], xyz[..., 2])
        cull = (alt <= 0) | (alt > dataset.config["ray_origin_height"])
        self.xyz, self.voxels = xyz[~cull], self.voxels[~cull]
        self.idx = torch.arange(self.xyz.shape[0], dtype=torch.int32)
        self.shp = (self.xyz.shape[0],)

    def dump(self, path: Path, sigma: torch.Tensor) -> None:
        """OpenVDB when its Python bindings are installed (harp2_extract.py:934-945), else the reference's
        own fallback (:919-933): the voxel indices and the extinction values as `voxels.npy` / `sigma.npy`,
        next to the requested file (the reference drops them into the working directory)."""
        path = Path(path)
        try:
            import openvdb as vdb  # noqa: PLC0415
        except ImportError:
            try:
                import pyopenvdb as vdb  # noqa: PLC0415
            except ImportError:
                vdb = None
        if vdb is None:
            voxel_file, sigma_file = path.parent / "voxels.npy", path.parent / "sigma.npy"
            if voxel_file.exists() or sigma_file.exists():
                raise FileExistsError(voxel_file if voxel_file.exists() else sigma_file)
            np.save(voxel_file, self.voxels.detach().cpu().numpy(), allow_pickle=False)
            np.save(sigma_file, sigma.detach().cpu().numpy(), allow_pickle=False)
            return
        if path.suffix != ".vdb":  # pragma: no cover - module absent here
            raise AssertionError("the global grid is written as an OpenVDB file")
        grid = vdb.FloatGrid()  # pragma: no cover
        values, voxels = sigma.detach().float().cpu().numpy(), self.voxels.cpu().numpy()
        for i in range(values.shape[0]):
            grid.copyFromArray(values[i, None, None, None], ijk=tuple(int(v) for v in voxels[i]))
        grid.transform = vdb.createLinearTransform(voxelSize=self.grid_res)
        grid.name, grid.saveFloatAsHalf, grid.vectorType = "density", True, "invariant"
        vdb.write(str(path), grids=[grid])
