"""Granule sources for HARP2Dataset: a real PACE/HARP2 netCDF file (when the `netCDF4` module
is installed) or a seeded synthetic HARP2-L1B-shaped granule (no files, no network).

Both expose the few raw fields HARP2Dataset consumes (reference: datasets/harp2.py:73-124,
:461-501): per-view angle and wavelength, and (V, H, W) arrays of latitude, longitude, surface
altitude, sensor zenith / azimuth and intensity, invalid values as NaN.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np

# HARP2 native view order: 10 blue (440 nm), 10 green (550), 60 red (670), 10 NIR (870)
HARP2_BANDS = ((440.0, 10), (550.0, 10), (670.0, 60), (870.0, 10))


class SyntheticGranule:
    """`synthetic:H=64,W=64,seed=0[,nan=0.02][,lat0=30][,lon0=-75]` -- geometry of a 5 x 5 degree scene
    (south-west corner lat0, lon0; default off the US east coast; longitudes wrap at the dateline) seen
    by 90 along-track views within +-45 degrees; intensities are a smooth function of
    position, band and view angle plus noise, with a fraction of NaN pixels."""

    processing_level = "L1B"

    def __init__(self, spec: str):
        opts = {"H": 64, "W": 64, "seed": 0, "nan": 0.02, "lat0": 30.0, "lon0": -75.0}
        body = spec.split(":", 1)[1] if ":" in spec else ""
        for item in filter(None, body.split(",")):
            k, v = item.split("=")
            opts[k] = float(v) if k in ("nan", "lat0", "lon0") else int(v)
        h, w = int(opts["H"]), int(opts["W"])
        rng = np.random.default_rng(int(opts["seed"]))
        wl, ang = [], []
        for nm, count in HARP2_BANDS:
            wl += [nm] * count
            ang += list(np.sort(rng.uniform(-45.0, 45.0, size=count)))
        self.wavelengths = np.asarray(wl, dtype=np.float32)
        self.view_angles = np.asarray(ang, dtype=np.float32)
        v = len(wl)
        # image rows run south -> north in the file (HARP2Dataset flips them so north is up)
        lat0, lon0 = float(opts["lat0"]), float(opts["lon0"])
        self.lat0, self.lon0, self.seed = lat0, lon0, int(opts["seed"])
        lat = np.linspace(lat0, lat0 + 5.0, h, dtype=np.float32)[:, None] + np.zeros((1, w), np.float32)
        lon = np.linspace(lon0, lon0 + 5.0, w, dtype=np.float32)[None, :] + np.zeros((h, 1), np.float32)
        lon = np.where(lon > 180.0, lon - 360.0, lon).astype(np.float32)
        self._fields = {
            "latitude": np.broadcast_to(lat, (v, h, w)).copy(),
            "longitude": np.broadcast_to(lon, (v, h, w)).copy(),
            "surface_altitude": np.zeros((v, h, w), np.float32),
        }
        a = self.view_angles[:, None, None]
        self._fields["sensor_zenith_angle"] = (np.abs(a) + rng.uniform(0, 0.2, (v, h, w))).astype(np.float32)
        self._fields["sensor_azimuth_angle"] = (np.where(a < 0, 180.0, 0.0) + rng.uniform(-2, 2, (v, h, w))).astype(np.float32)
        yy, xx = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
        base = 0.08 + 0.05 * np.sin(6 * xx)[None] * np.cos(4 * yy)[None]
        band_gain = (self.wavelengths[:, None, None] / 870.0) ** -1.5 * 0.35
        limb = 1.0 / np.cos(np.deg2rad(a))
        img = (base * band_gain * limb + rng.uniform(0, 0.01, (v, h, w))).astype(np.float32)
        img[rng.random((v, h, w)) < float(opts["nan"])] = np.nan
        self._fields["i"] = np.minimum(img, 0.3).astype(np.float32)

    def field(self, name: str) -> np.ndarray:
        return self._fields[name]

    # ---- stand-ins for the two auxiliary products the extract layouts read (datasets/harp2_extract.py) ----
    def l1c_geolocation(self, step_km: float = 5.0) -> dict[str, np.ndarray]:
        """What `HARP2L1CExtractDataset` reads from the granule's L1C file (harp2_extract.py:149-166): the
        5 km L1C bin grid of the scene, float32 (bins_along_track, bins_across_track) latitude, longitude
        and height, rows south -> north like the file (the reader flips them), fill values as NaN (the two
        southern corner bins, as in swath-shaped L1C grids)."""
        step = step_km / 111.0
        n_lat, n_lon = int(5.0 / step), int(5.0 / (step / np.cos(np.deg2rad(self.lat0 + 2.5))))
        lat = (self.lat0 + (np.arange(n_lat, dtype=np.float32) + 0.5) * np.float32(5.0 / n_lat))[:, None] + np.zeros((1, n_lon), np.float32)
        lon = (self.lon0 + (np.arange(n_lon, dtype=np.float32) + 0.5) * np.float32(5.0 / n_lon))[None, :] + np.zeros((n_lat, 1), np.float32)
        lon = np.where(lon > 180.0, lon - 360.0, lon).astype(np.float32)
        rng = np.random.default_rng(self.seed + 101)
        height = (rng.uniform(0.0, 40.0, (n_lat, n_lon))).astype(np.float32)
        out = {"latitude": lat.astype(np.float32).copy(), "longitude": lon.copy(), "height": height}
        for a in out.values():
            a[0, 0] = a[0, -1] = np.nan
        return out

    def earthcare_track(self, n_along: int = 160, n_height: int = 90) -> dict[str, np.ndarray]:
        """What `HARP2EarthCAREExtractDataset` reads from an ATLID `ATL_EBD_2A` file
        (harp2_extract.py:626-650): a ground track crossing the scene from its south-west to its
        north-east corner (float64 latitude / longitude per profile) and the joint-standard-grid height of
        every range bin, top first, from above the shell down to below the ellipsoid, slightly different
        from profile to profile (so the reference's all-profiles altitude mask has something to drop)."""
        t = np.linspace(-0.1, 1.1, n_along)
        lat = self.lat0 + 5.0 * t
        lon = self.lon0 + 5.0 * t
        lon = np.where(lon > 180.0, lon - 360.0, lon)
        rng = np.random.default_rng(self.seed + 202)
        base = np.linspace(24000.0, -600.0, n_height)[None, :]
        height = base + rng.uniform(-40.0, 40.0, (n_along, 1)) + np.zeros((1, n_height))
        return {"latitude": lat.astype(np.float64), "longitude": lon.astype(np.float64),
                "height": height.astype(np.float64), "file_type": "ATL_EBD_2A"}


class NetCDFGranule:
    """A HARP2 L1B / L1C file read with netCDF4 (same variables as datasets/harp2.py:100-110)."""

    def __init__(self, path: Path):
        try:
            import netCDF4  # noqa: PLC0415
        except ImportError as err:  # pragma: no cover - module absent in the build container
            raise ImportError("reading HARP2 netCDF granules needs the `netCDF4` module") from err
        self.nc = netCDF4.Dataset(path)
        self.processing_level = self.nc.processing_level
        if self.processing_level not in ("L1B", "L1C"):
            raise NotImplementedError(f"HARP2 level {self.processing_level}")
        self.view_angles = self.nc["sensor_views_bands/sensor_view_angle"][:].filled(fill_value=np.nan)
        self.wavelengths = self.nc["sensor_views_bands/intensity_wavelength"][:].data.flatten()

    _PATHS = {
        "latitude": "geolocation_data/latitude", "longitude": "geolocation_data/longitude",
        "surface_altitude": "geolocation_data/surface_altitude", "height": "geolocation_data/height",
        "sensor_zenith_angle": "geolocation_data/sensor_zenith_angle",
        "sensor_azimuth_angle": "geolocation_data/sensor_azimuth_angle", "i": "observation_data/i",
    }

    def field(self, name: str) -> np.ndarray:
        if name == "surface_altitude" and self.processing_level == "L1C":
            name = "height"
        return self.nc[self._PATHS[name]][:].filled(fill_value=np.nan)


def open_granule(filename: str, local_dir: Path):
    if filename.startswith("synthetic"):
        return SyntheticGranule(filename)
    return NetCDFGranule(local_dir / filename)
