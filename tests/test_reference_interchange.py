"""Checkpoint interchange with the reference's own NeRFPipeline (the Instant-NGP pipeline of the
reference cannot be constructed here: tiny-cuda-nn is absent). The reference package is also called
`atmonr`, so it runs in a child process (tests/golden/make_golden.py: import_reference). Only
possible in the build container; skipped where /root/reference does not exist (the GPU box)."""

import json
import os
import subprocess
import sys

import pytest
import torch

from helpers import ROOT, FakeDataset, tiny_scene

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")

CHILD = r"""
import sys, json, torch
from types import SimpleNamespace
sys.path.insert(0, sys.argv[1])
from make_golden import import_reference
import_reference()
from atmonr.pipelines.nerf import NeRFPipeline
cfg = json.loads(sys.argv[2])
ds = SimpleNamespace(config={"ray_origin_height": 20000}, scale=1.0, offset=torch.zeros(3, dtype=torch.float64),
                     max_i=0.3, get_point_preprocessor=lambda name: (lambda p: p))
torch.manual_seed(3)
pipe = NeRFPipeline(cfg, ds)
if sys.argv[3] == "save":
    torch.save(pipe.state_dict(), sys.argv[4])
else:
    pipe.load_state_dict(torch.load(sys.argv[4]))      # strict: every key and shape must match
    sd = pipe.state_dict()
    torch.save({m: {k: v.clone() for k, v in sd[m].items()} for m in sd}, sys.argv[5])
"""


def _child(*args):
    r = subprocess.run([sys.executable, "-c", CHILD, os.path.join(ROOT, "tests", "golden"), *args],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]


def test_nerf_checkpoint_interchanges_with_the_reference(tmp_path):
    from atmonr.pipelines.factory import get_pipeline
    cfg = json.load(open(os.path.join(ROOT, "configs", "nerf.json")))["pipeline"]
    cfg["mlp_hidden_dim"] = 32
    ref_pt, mine_pt, back_pt = (str(tmp_path / n) for n in ("ref.pt", "mine.pt", "back.pt"))
    # reference -> this package
    _child(json.dumps(cfg), "save", ref_pt)
    ref_sd = torch.load(ref_pt)
    pipe = get_pipeline(cfg, FakeDataset(tiny_scene(h=2, w=2, n_views=3)))
    pipe.load_state_dict(ref_sd)
    mine = pipe.state_dict()
    assert set(mine) == set(ref_sd) == {"coarse", "fine"}
    for mode in mine:
        assert list(mine[mode]) == list(ref_sd[mode])                  # same names, same order (fc1..fc11)
        for k in mine[mode]:
            assert torch.equal(mine[mode][k], ref_sd[mode][k]), (mode, k)
    # this package -> reference
    torch.manual_seed(8)
    pipe2 = get_pipeline(cfg, FakeDataset(tiny_scene(h=2, w=2, n_views=3)))
    torch.save(pipe2.state_dict(), mine_pt)
    _child(json.dumps(cfg), "load", mine_pt, back_pt)
    back = torch.load(back_pt)
    for mode in back:
        for k, v in pipe2.state_dict()[mode].items():
            assert torch.equal(back[mode][k], v), (mode, k)


# ------------------------------------------------------------------------------------------
# the reference's HARP2Dataset on the synthetic granule
# ------------------------------------------------------------------------------------------
DATASET_CHILD = r"""
import importlib.util, json, os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, sys.argv[1])
from make_golden import import_reference
import_reference()
torch.Tensor.cuda = lambda self, *a, **k: self            # the reference moves every array to the GPU
spec = importlib.util.spec_from_file_location("granule_src", sys.argv[2])   # this repository's granule source
gmod = importlib.util.module_from_spec(spec); spec.loader.exec_module(gmod)
gran = gmod.SyntheticGranule(sys.argv[3])

class Var:
    def __init__(self, a): self.a = np.ma.MaskedArray(a, mask=np.isnan(a)); self.shape = a.shape
    def __getitem__(self, k): return self.a[k]

class FakeNC:
    processing_level = "L1B"
    def __init__(self, path): pass
    def __getitem__(self, path):
        group, name = path.split("/")
        if name == "sensor_view_angle": return Var(gran.view_angles)
        if name == "intensity_wavelength": return Var(gran.wavelengths[None])
        return Var(gran.field(name))

import netCDF4
netCDF4.Dataset = FakeNC
os.chdir(sys.argv[5])
Path("data/HARP2").mkdir(parents=True)
Path("data/HARP2/fake.nc").touch()
from atmonr.datasets.harp2 import HARP2Dataset
ds = HARP2Dataset(json.loads(sys.argv[4]), "fake.nc")
idx = torch.arange(0, len(ds), 97)
batch = ds[idx]
pre = ds.get_point_preprocessor("horizontal")
g = torch.Generator().manual_seed(0)
p32 = torch.rand(2, 50, 3, generator=g) * 1.6 - 0.8
p64 = (torch.rand(1, 50, 3, generator=g, dtype=torch.float64) * 1.6 - 0.8)
torch.save({"view_idx": ds.view_idx, "irgb_idx": ds.irgb_idx, "img_shp": tuple(ds.img_shp), "max_i": ds.max_i,
            "ray_filter": ds.ray_filter, "ray_origin_norm": ds.ray_origin_norm, "ray_dir": ds.ray_dir,
            "ray_len_norm": ds.ray_len_norm, "ray_rad": ds.ray_rad, "ray_alt": ds.ray_alt,
            "ray_irgb_idx": ds.ray_irgb_idx, "scale": ds.scale, "offset": ds.offset, "len": len(ds),
            "batch": {k: v for k, v in batch.items()}, "idx": idx, "p32": p32, "p64": p64,
            "pre32": pre(p32), "pre64": pre(p64), "best_rgb_idx": [int(v) for v in ds.best_rgb_idx],
            "tracker": {k: getattr(ds.get_progress_tracker(), k) for k in
                        ("valid", "target_img", "target_img_rgb", "pred_img", "pred_pixels")}}, sys.argv[6])
"""


def test_dataset_matches_the_reference_dataset_on_the_synthetic_granule(tmp_path):
    """a1 + a2 + a3 + the `horizontal` closure through the reference's OWN HARP2Dataset
    (datasets/harp2.py:26-429), fed the synthetic granule through a stand-in for netCDF4.Dataset."""
    from atmonr.datasets.factory import get_dataset
    from atmonr.datasets.harp2 import HARP2Dataset
    spec = "synthetic:H=10,W=9,seed=4"
    cfg = json.load(open(os.path.join(ROOT, "configs", "instant_ngp.json")))["dataset"]
    out = str(tmp_path / "ref_ds.pt")
    r = subprocess.run([sys.executable, "-c", DATASET_CHILD, os.path.join(ROOT, "tests", "golden"),
                        os.path.join(ROOT, "atmospheric-neural-rendering_b200", "atmonr", "datasets", "granule.py"),
                        spec, json.dumps(cfg), str(tmp_path), out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = torch.load(out, weights_only=False)
    ds = HARP2Dataset(dict(cfg), spec, device=torch.device("cpu"))
    assert (ds.view_idx == ref["view_idx"]).all() and (ds.irgb_idx == ref["irgb_idx"]).all()
    assert tuple(ds.img_shp) == ref["img_shp"] and len(ds) == ref["len"]
    assert abs(ds.max_i - ref["max_i"]) <= 1e-7
    assert torch.equal(ds.ray_filter, ref["ray_filter"])
    assert ds.scale == ref["scale"] and torch.equal(ds.offset, ref["offset"])
    for k in ("ray_origin_norm", "ray_dir", "ray_len_norm", "ray_rad", "ray_alt", "ray_irgb_idx"):
        assert torch.equal(getattr(ds, k), ref[k]), k          # same torch build, same operations: bit for bit
    # visualisation side of the dataset (harp2.py:126-203, 259-349): RGB view choice and progress tracker
    import numpy as np
    assert [int(v) for v in ds.best_rgb_idx] == ref["best_rgb_idx"]
    tracker = ds.get_progress_tracker()
    for k, v in ref["tracker"].items():
        got = getattr(tracker, k)
        assert got.shape == v.shape and np.array_equal(got, v, equal_nan=True), k
    batch = ds[ref["idx"]]
    assert set(batch) == set(ref["batch"])
    for k, v in ref["batch"].items():
        assert batch[k].dtype == v.dtype and torch.equal(batch[k], v), k
    # the preprocessor closure: the oracle (pinned to the reference elsewhere) built from this dataset's frame
    from oracle import geodesy
    lat, lon = ds.lat[~ds.lat.isnan()], ds.lon[~ds.lon.isnan()]
    frame = geodesy.HorizontalFrame.from_latlon(lat, lon, ds.scale, ds.offset, 20000.0)
    assert torch.equal(geodesy.preprocess_horizontal(ref["p32"], frame), ref["pre32"])
    assert float((geodesy.preprocess_horizontal(ref["p64"], frame) - ref["pre64"]).abs().max()) <= 1e-12
