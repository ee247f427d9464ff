"""HARP2 dataset (reference: src/atmonr/datasets/harp2.py): turns a granule into the ray tables
the pipelines consume, with the same attribute surface as the reference's HARP2Dataset
(`lat, lon, alt, img_shp, view_idx, irgb_idx, max_i, ray_* tables, scale, offset, ray_filter,
best_rgb_idx, get_rgb, get_image_metrics, get_progress_tracker, get_point_preprocessor,
__getitem__/__getbatch__/__len__`).

The granule comes from a netCDF file (if `netCDF4` is installed) or from the built-in synthetic
generator (`--scene-filename synthetic:H=64,W=64,seed=0`), see datasets/granule.py.
"""

from __future__ import annotations

import os
from pathlib import Path
from typing import Callable

import numpy as np
import torch
from torch.utils.data import Dataset

from atmonr.datasets.granule import open_granule
from atmonr.geospatial.wgs_84 import filter_rays, get_rays, normalize_rays
from atmonr.native import lib as L
from atmonr.native import ops
from atmonr.progress_tracker import ProgressTracker


def get_indexes(view_angles, wavelengths, max_abs_view_angle: float, bands_to_keep=(0, 1, 2, 3)):
    """harp2.py:461-501: views within the angle limit, re-ordered by decreasing wavelength
    (IRGB), and each kept view's band index (0 NIR, 1 red, 2 green, 3 blue)."""
    keep = np.where(np.abs(view_angles) <= max_abs_view_angle)[0]
    order = np.argsort(-wavelengths, stable=True)
    view_idx = order[np.isin(order, keep)]
    irgb_idx = np.where(wavelengths[view_idx, None] == np.unique(wavelengths)[None, ::-1])[1]
    mask = np.isin(irgb_idx, list(bands_to_keep))
    return view_idx[mask], irgb_idx[mask]


class HorizontalPreprocessor:
    """The 'horizontal' point preprocessor (harp2.py:351-388) as a callable object. `frame`
    carries the closure's constants for the fused kernels; calling it runs the standalone
    kernel (float32 or float64 points of shape (..., 3))."""

    def __init__(self, frame: L.FrameT):
        self.frame = frame

    def __call__(self, coords_xyz: torch.Tensor) -> torch.Tensor:
        return ops.preprocess_horizontal(self.frame, coords_xyz)


class HARP2Dataset(Dataset):
    def __init__(self, config: dict, filename: str, chunk_size: int = int(1e4), device=None) -> None:
        super().__init__()
        self.config = config
        self.filename = filename
        self.local_path = Path("data/HARP2") / filename
        self.config.setdefault("max_abs_view_angle", 90.0)
        self.device = device or (torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu"))
        self.granule = open_granule(filename, Path("data/HARP2"))
        self.view_idx, self.irgb_idx = get_indexes(
            self.granule.view_angles, self.granule.wavelengths,
            self.config["max_abs_view_angle"], self.config.get("bands_to_keep", [0, 1, 2, 3]),
        )
        with torch.no_grad():
            self._init_data()
            # configs/nerf.json has no rgb_mode (reference would KeyError)
            self._init_rgb_idxs(self.config.get("rgb_mode", "nadir"))
            self._init_ray_data(chunk_size)

    # ---- harp2.py:73-124 -------------------------------------------------------------------
    def _parse(self, name: str) -> np.ndarray:
        arr = self.granule.field(name)
        nv = self.view_idx.shape[0]
        if self.granule.processing_level == "L1B":  # (V, H, W): keep views, north up, angle last
            return arr[self.view_idx, ::-1].transpose((1, 2, 0)).reshape((-1, nv))
        if arr.ndim == 4:
            arr = arr[..., 0]
        if arr.ndim == 3:
            return arr[::-1, :, self.view_idx].reshape((-1, nv))
        return np.tile(arr[::-1, :, None], (1, 1, nv)).reshape((-1, nv))

    def _init_data(self) -> None:
        raw_i = self.granule.field("i")
        self.img_shp = raw_i.shape[1:] if self.granule.processing_level == "L1B" else raw_i.shape[:2]
        to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)
        self.lat, self.lon = to_dev(self._parse("latitude")), to_dev(self._parse("longitude"))
        self.alt = to_dev(self._parse("surface_altitude"))
        self.thetav = to_dev(self._parse("sensor_zenith_angle"))
        self.phiv = to_dev(self._parse("sensor_azimuth_angle"))
        i = self._parse("i")
        self.max_i = np.nanmax(i).item()
        self.int_arr = to_dev(i)

    # ---- harp2.py:126-198 ------------------------------------------------------------------
    def _init_rgb_idxs(self, mode: str = "nadir") -> None:
        angles = np.asarray(self.granule.view_angles)[self.view_idx]
        num_valid = (~self.int_arr.isnan()).sum(dim=0).cpu().numpy()
        striped = np.zeros_like(num_valid, dtype=bool)
        if self.granule.processing_level == "L1B":
            striped = num_valid < num_valid.mean()
        masks = [self.irgb_idx == b for b in (1, 2, 3)]
        idxs = [np.where(m)[0] for m in masks]
        ang = [angles[m] for m in masks]
        if not masks[0].any():
            best = int(np.argmin(np.abs(angles) + striped * 1000))
            self.best_rgb_idx = [best] * 3
            return
        if not masks[1].any() or not masks[2].any():
            best = int(idxs[0][int(np.argmin(np.abs(ang[0]) + striped[masks[0]] * 1000))])
            self.best_rgb_idx = [best] * 3
            return
        mesh = np.stack(np.meshgrid(*ang, indexing="ij"))
        spread = (mesh.max(axis=0) - mesh.min(axis=0)).reshape((ang[0].shape[0], -1))
        nearest = spread.argmin(axis=1)
        green = idxs[1][nearest // ang[2].shape[0]]
        blue = idxs[2][nearest % ang[2].shape[0]]
        if mode == "nadir":
            pick = int(np.argmin(np.abs(ang[0]) + striped[masks[0]] * 1000))
        elif mode == "most_pixels":
            pick = int(np.stack([num_valid[masks[0]], num_valid[green], num_valid[blue]]).min(axis=0).argmax(axis=0))
        else:
            raise NotImplementedError(f"Unrecognized RGB indexing mode {mode}")
        self.best_rgb_idx = [int(idxs[0][pick]), int(green[pick]), int(blue[pick])]

    # ---- harp2.py:200-257 ------------------------------------------------------------------
    def _init_ray_data(self, chunk_size: int) -> None:
        n_pix, n_view = self.lat.shape
        origins, dirs, lens = [], [], []
        for lo in range(0, n_pix, chunk_size):
            sl = slice(lo, min(lo + chunk_size, n_pix))
            o, d, ln = get_rays(self.lat[sl], self.lon[sl], self.alt[sl], self.thetav[sl], self.phiv[sl],
                                ray_origin_height=self.config["ray_origin_height"])
            origins.append(o), dirs.append(d), lens.append(ln)
        ray_origin, ray_dir, ray_len = torch.cat(origins), torch.cat(dirs), torch.cat(lens)
        ray_rad = self.int_arr.flatten()
        self.ray_filter = filter_rays(ray_origin, ray_dir, ray_rad)
        keep = self.ray_filter
        self.ray_origin, self.ray_dir = ray_origin[keep], ray_dir[keep].contiguous()
        self.ray_rad, self.ray_len = ray_rad[keep], ray_len[keep]
        self.ray_alt = self.alt.flatten()[keep]
        self.ray_origin_norm, self.scale, self.offset = normalize_rays(self.ray_origin, self.ray_dir, self.ray_len)
        self.ray_len_norm = self.ray_len / self.scale
        band_of_view = torch.from_numpy(self.irgb_idx).to(device=keep.device)
        self.ray_irgb_idx = band_of_view[torch.where(keep.view((-1, n_view)))[1]]
        # (the reference keeps ray_idx on the host, harp2.py:254, which costs a blocking copy per batch;
        # here it lives with the ray tables, and the trainer scatters with it on the device)
        self.ray_idx = torch.arange(self.ray_origin_norm.shape[0], dtype=torch.int32, device=self.ray_origin_norm.device)

    # ---- harp2.py:259-349 ------------------------------------------------------------------
    def get_progress_tracker(self) -> ProgressTracker:
        nv = self.view_idx.shape[0]
        target = torch.zeros(self.img_shp[0] * self.img_shp[1] * nv, device=self.ray_filter.device)
        target[self.ray_filter] = self.ray_rad
        target = target.view(list(self.img_shp) + [nv])
        zeros_img = lambda: np.zeros(target.shape, dtype=np.float32)
        zeros_pix = lambda: np.zeros(self.ray_rad.shape, dtype=np.float32)
        return ProgressTracker(
            valid=self.ray_filter.view(self.img_shp[0], self.img_shp[1], nv).cpu().numpy(),
            target_img=target.cpu().numpy(),
            target_img_rgb=self.get_rgb(target.permute((2, 0, 1))).cpu().numpy(),
            pred_img=zeros_img(), pred_pixels=zeros_pix(),
            pred_img_surf=zeros_img(), pred_pixels_surf=zeros_pix(),
            pred_img_atmo=zeros_img(), pred_pixels_atmo=zeros_pix(),
        )

    def get_image_metrics(self, pred_img: torch.Tensor, target_img: torch.Tensor) -> dict:
        """PSNR / SSIM per view (V, H, W inputs). torchmetrics is not a dependency here: both are
        evaluated directly (SSIM: 11x11 Gaussian window, sigma 1.5, k1 0.01, k2 0.03)."""
        pred = torch.clip(pred_img / self.max_i, min=0, max=1)
        target = target_img / self.max_i
        data_range = (target.max() - target.min()).item()
        mse = ((pred - target) ** 2).mean(dim=(1, 2))
        psnr = 10 * torch.log10(data_range**2 / mse)
        ssim = _ssim(pred[:, None], target[:, None], data_range)
        return {
            "PSNR": psnr.cpu().numpy().tolist(), "SSIM": ssim.cpu().numpy().tolist(),
            "PSNR_mean": psnr[~torch.isnan(psnr)].mean().item(),
            "SSIM_mean": ssim[~torch.isnan(ssim)].mean().item(),
        }

    def get_rgb(self, cube: torch.Tensor) -> torch.Tensor:
        assert cube.shape == (self.view_idx.shape[0], self.img_shp[0], self.img_shp[1])
        return torch.clamp(cube[self.best_rgb_idx] / self.max_i, 0, 1).permute(1, 2, 0).contiguous()

    # ---- harp2.py:351-390 ------------------------------------------------------------------
    def get_point_preprocessor(self, point_preprocessor: str) -> Callable:
        if point_preprocessor != "horizontal":
            raise NotImplementedError
        lat = self.lat[~self.lat.isnan()]
        lon = self.lon[~self.lon.isnan()]
        lat_min, lat_max, lon_min, lon_max = lat.min(), lat.max(), lon.min(), lon.max()
        shift_lon = bool(lon_max > 179 and lon_min < -179)
        if shift_lon:  # the granule straddles the dateline
            lon = lon % 360 - 180
            lon_min, lon_max = lon.min(), lon.max()
        frame = L.make_frame(self.scale, self.offset.tolist(), lat_min.item(), (lat_max - lat_min).item(),
                             lon_min.item(), (lon_max - lon_min).item(), self.config["ray_origin_height"], shift_lon)
        return HorizontalPreprocessor(frame)

    # ---- harp2.py:392-429 ------------------------------------------------------------------
    def __getitem__(self, idx) -> dict[str, torch.Tensor]:
        return {
            "origin": self.ray_origin_norm[idx], "dir": self.ray_dir[idx], "alt": self.ray_alt[idx],
            "rad": self.ray_rad[idx], "len": self.ray_len_norm[idx], "idx": self.ray_idx[idx],
            "irgb_idx": self.ray_irgb_idx[idx],
        }

    def __getbatch__(self, idx: torch.Tensor) -> dict[str, torch.Tensor]:
        # the seven gathers in one launch (atmonr_gather_batch, csrc/rays.cu); ATMONR_NATIVE_GATHER=0 keeps
        # torch indexing (cross-check)
        if os.environ.get("ATMONR_NATIVE_GATHER", "1") != "0" and self.ray_origin_norm.is_cuda and torch.is_tensor(idx) \
                and idx.dtype == torch.int64 and idx.dim() == 1:
            return ops.gather_batch(self._ray_tables(), idx.to(self.ray_origin_norm.device))
        return self[idx]

    def _ray_tables(self) -> dict[str, torch.Tensor]:
        return {"origin": self.ray_origin_norm, "dir": self.ray_dir, "alt": self.ray_alt, "rad": self.ray_rad,
                "len": self.ray_len_norm, "idx": self.ray_idx, "irgb_idx": self.ray_irgb_idx}

    def __len__(self) -> int:
        return self.ray_origin_norm.shape[0]


def _ssim(x: torch.Tensor, y: torch.Tensor, data_range: float, size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    coords = torch.arange(size, dtype=x.dtype, device=x.device) - (size - 1) / 2
    g = torch.exp(-(coords**2) / (2 * sigma**2))
    g = (g / g.sum())[:, None] * (g / g.sum())[None, :]
    win = g[None, None]
    pad = size // 2
    xp = torch.nn.functional.pad(x, (pad, pad, pad, pad), mode="reflect")
    yp = torch.nn.functional.pad(y, (pad, pad, pad, pad), mode="reflect")
    conv = lambda t: torch.nn.functional.conv2d(t, win)
    mu_x, mu_y = conv(xp), conv(yp)
    sxx, syy, sxy = conv(xp * xp) - mu_x**2, conv(yp * yp) - mu_y**2, conv(xp * yp) - mu_x * mu_y
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    ssim_map = ((2 * mu_x * mu_y + c1) * (2 * sxy + c2)) / ((mu_x**2 + mu_y**2 + c1) * (sxx + syy + c2))
    return ssim_map.mean(dim=(1, 2, 3))
