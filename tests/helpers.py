"""Shared fixtures for the parity tests: a tiny synthetic scene built with the ORACLE's geodesy,
oracle parameters, and loaders that push them into the native pipeline."""

from __future__ import annotations

import json
import os
from types import SimpleNamespace

import numpy as np
import torch

from oracle import geodesy
from oracle.ngp import NGPOracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngp_config(n_samples=None):
    cfg = json.load(open(os.path.join(ROOT, "configs", "instant_ngp.json")))["pipeline"]
    if n_samples:
        cfg["num_samples_per_ray"] = n_samples
    return cfg


def tiny_scene(h=8, w=8, n_views=9, seed=0):
    """HARP2-shaped geometry -> normalised rays (oracle geodesy, CPU)."""
    rng = np.random.default_rng(seed)
    p = h * w
    lat = np.repeat(np.linspace(35.0, 30.0, h, dtype=np.float32)[:, None], w, 1).reshape(p, 1) + np.zeros((1, n_views), np.float32)
    lon = np.repeat(np.linspace(-75.0, -70.0, w, dtype=np.float32)[None, :], h, 0).reshape(p, 1) + np.zeros((1, n_views), np.float32)
    alt = np.zeros((p, n_views), np.float32)
    ang = np.linspace(-44.0, 44.0, n_views, dtype=np.float32)
    thetav = (np.abs(ang)[None, :] + rng.random((p, n_views)) * 0.2).astype(np.float32)
    phiv = (np.where(ang[None, :] < 0, 180.0, 0.0) + rng.random((p, n_views)) * 4 - 2).astype(np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    lat, lon, alt, thetav, phiv = map(t, (lat, lon, alt, thetav, phiv))
    o, d, ln = geodesy.build_rays(lat, lon, alt, thetav, phiv, 20000.0)
    on, scale, offset = geodesy.normalize_rays(o, d, ln)
    frame = geodesy.HorizontalFrame.from_latlon(lat, lon, scale, offset, 20000.0)
    g = torch.Generator().manual_seed(seed)
    n = on.shape[0]
    batch = {
        "origin": on, "dir": d, "len": ln / scale,
        "rad": torch.rand(n, generator=g) * 0.3, "irgb_idx": torch.randint(0, 4, (n,), generator=g),
        "idx": torch.arange(n, dtype=torch.int32),
    }
    return SimpleNamespace(lat=lat, lon=lon, scale=scale, offset=offset, frame=frame, batch=batch, max_i=0.3)


def take(batch, sl):
    return {k: v[sl].contiguous() for k, v in batch.items()}


def to_cuda(batch):
    return {k: (v if k == "idx" else v.cuda()) for k, v in batch.items()}


class FakeDataset:
    """Duck-typed dataset for constructing native pipelines in tests (what Pipeline.__init__ reads)."""

    def __init__(self, scene):
        from atmonr.datasets.harp2 import HorizontalPreprocessor
        from atmonr.native import lib as L

        self.config = {"ray_origin_height": 20000, "type": "HARP2"}
        self.scale, self.offset, self.max_i = scene.scale, scene.offset, scene.max_i
        fr = scene.frame
        self._pre = HorizontalPreprocessor(L.make_frame(fr.scale, fr.offset, fr.lat_min, fr.lat_range, fr.lon_min,
                                                        fr.lon_range, fr.origin_height, fr.shift_lon))

    def get_point_preprocessor(self, name):
        assert name == "horizontal"
        return self._pre


def random_params(oracle: NGPOracle, seed=0, table_scale=1.0):
    """Oracle-initialised parameters, with larger table values than tcnn's U(-1e-4,1e-4) init so
    that the outputs are not numerically trivial."""
    params = oracle.init_params(seed)
    with torch.no_grad():
        for k in ("pos_encoder", "surf_encoder"):
            params[k].mul_(table_scale)
    return params


def load_params(pipeline, params):
    pipeline.load_state_dict({k: {"params": v.detach().clone()} for k, v in params.items()})
