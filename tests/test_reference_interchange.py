"""Checkpoint interchange with the reference's own NeRFPipeline (the Instant-NGP pipeline of the
reference cannot be constructed here: tiny-cuda-nn is absent). The reference package is also called
`atmonr`, so it runs in a child process (tests/golden/make_golden.py: import_reference). Only
possible in the build container; skipped where /root/reference does not exist (the GPU box)."""

import json
import os
import subprocess
import sys

import numpy as np
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


@pytest.mark.parametrize("variant", ["shipped", "most_pixels_rgb_30deg_no_nir", "dateline"])
def test_dataset_matches_the_reference_dataset_on_the_synthetic_granule(tmp_path, variant):
    """a1 + a2 + a3 + the `horizontal` closure through the reference's OWN HARP2Dataset
    (datasets/harp2.py:26-429), fed the synthetic granule through a stand-in for netCDF4.Dataset."""
    from atmonr.datasets.factory import get_dataset
    from atmonr.datasets.harp2 import HARP2Dataset
    spec = "synthetic:H=10,W=9,seed=4"
    cfg = json.load(open(os.path.join(ROOT, "configs", "instant_ngp.json")))["dataset"]
    if variant == "dateline":     # a granule across the dateline: the lon-shift branch of harp2.py:366-370
        spec = "synthetic:H=10,W=9,seed=4,lat0=-62,lon0=177.5"
    elif variant != "shipped":    # the other RGB view choice, a tighter view-angle filter, a dropped band
        cfg.update(rgb_mode="most_pixels", max_abs_view_angle=30.0, bands_to_keep=[1, 2, 3])
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
    assert frame.shift_lon == (variant == "dateline")
    # the constants this package hands to the sampler kernel are the closure's
    fr = ds.get_point_preprocessor("horizontal").frame
    assert bool(fr.shift_lon) == frame.shift_lon and fr.enabled == 1
    for name in ("scale", "lat_min", "lat_range", "lon_min", "lon_range", "origin_height"):
        assert getattr(fr, name) == getattr(frame, name), name
    assert tuple(fr.offset) == tuple(frame.offset)
    assert torch.equal(geodesy.preprocess_horizontal(ref["p32"], frame), ref["pre32"])
    assert float((geodesy.preprocess_horizontal(ref["p64"], frame) - ref["pre64"]).abs().max()) <= 1e-12


# ------------------------------------------------------------------------------------------
# the reference's InstantNGPPipeline around a stand-in tinycudann
# ------------------------------------------------------------------------------------------
NGP_CHILD = DATASET_CHILD.split("from atmonr.datasets.harp2 import HARP2Dataset")[0] + r"""
# a stand-in `tinycudann`: the module API the reference uses (instant_ngp.py:60-85), with the
# arithmetic of oracle/tcnn_spec.py in float32. What is under test is everything AROUND it.
import types
sys.path.insert(0, sys.argv[7])                      # repository root: the `oracle` package
from oracle import tcnn_spec

class _Mod(torch.nn.Module):
    def __init__(self, impl):
        super().__init__()
        self.impl, self.n_output_dims = impl, impl.n_output_dims
        self.params = torch.nn.Parameter(torch.zeros(impl.n_params))
    def forward(self, x):
        return self.impl.forward(x.float(), self.params, False)

tcnn = types.ModuleType("tinycudann")
tcnn.Encoding = lambda n_in, cfg, **kw: _Mod(tcnn_spec.make_encoding(n_in, cfg))
tcnn.Network = lambda n_in, n_out, cfg, **kw: _Mod(tcnn_spec.Network(n_in, n_out, cfg))
sys.modules["tinycudann"] = tcnn

from atmonr.datasets.harp2 import HARP2Dataset
from atmonr.pipelines.instant_ngp import InstantNGPPipeline
cfg = json.loads(sys.argv[4])
ds = HARP2Dataset(cfg["dataset"], "fake.nc")
pipe = InstantNGPPipeline(cfg["pipeline"], ds)
pipe.send_tensors_to(0)
job = torch.load(sys.argv[8])
for name, p in job["params"].items():
    getattr(pipe, name).params.data.copy_(p)
batch = ds[job["idx"]]
torch.manual_seed(job["seed"])
res = pipe.forward(batch)
loss = pipe.compute_loss(batch, res)
loss.backward()
pipe.eval()
sig = pipe.extract(job["pts"].clone())
opt = pipe.get_optimizer(cfg["trainer"]["optimizer"])
curve = []
if job.get("steps"):                                  # trainer.py:99-105, the reference's own step order
    pipe.train()
    for k in range(job["steps"]):
        b = ds[job["idx"] + k]
        torch.manual_seed(job["seed"] + 1 + k)
        r_ = pipe.forward(b)
        l_ = pipe.compute_loss(b, r_)
        opt.zero_grad()
        l_.backward()
        opt.step()
        curve.append(float(l_.detach()))
torch.save({"curve": curve, "res": {k: v.detach() for k, v in res.items()}, "loss": loss.detach(),
            "grads": {n: getattr(pipe, n).params.grad for n in job["params"] if getattr(pipe, n).params.grad is not None},
            "extract": sig.detach(), "state_keys": {k: list(v) for k, v in pipe.state_dict().items()},
            "groups": [(len(g["params"]), g["weight_decay"], g["lr"], tuple(g["betas"]), g["eps"]) for g in opt.param_groups]},
           sys.argv[6])
"""


@pytest.mark.parametrize("variant", ["default", "multi_band_extinction", "include_height", "l1_plus_hdr"])
def test_ngp_glue_matches_the_reference_pipeline(tmp_path, variant):
    """instant_ngp.py:129-263 (sampling, preprocessing, remap, altitude compression, direction
    conditioning, ReLUs, compositing with the surface, z in km, band selection and loss, extract) run
    from the REFERENCE's own InstantNGPPipeline, with a stand-in for the absent tiny-cuda-nn that
    evaluates oracle/tcnn_spec.py. The oracle pipeline (oracle/ngp.py) must reproduce its results and
    parameter gradients: this pins the oracle's glue; only the tcnn arithmetic itself stays unpinned."""
    from atmonr.datasets.harp2 import HARP2Dataset
    from oracle import geodesy
    from oracle.ngp import NGPOracle
    spec = "synthetic:H=10,W=9,seed=4"
    cfg = json.load(open(os.path.join(ROOT, "configs", "instant_ngp.json")))
    cfg["pipeline"]["num_samples_per_ray"] = 24
    for key in ("encoding", "surface_encoding"):
        cfg["pipeline"]["instant_ngp"][key]["log2_hashmap_size"] = 12       # small tables: CPU-sized test
    if variant == "multi_band_extinction":                                  # one density per band
        cfg["pipeline"]["multi_band_extinction"] = True
    elif variant == "include_height":                                       # 4-D grid on raw scene coordinates
        cfg["pipeline"]["include_height"], cfg["pipeline"]["point_preprocessor"] = True, ""
        cfg["pipeline"]["encoder"] = {"L_x": 4}       # pipeline.py:36 reads this NeRF key whenever there is no preprocessor
    elif variant == "l1_plus_hdr":
        cfg["pipeline"]["loss"] = "l1_plus_hdr"
    ds = HARP2Dataset(dict(cfg["dataset"]), spec, device=torch.device("cpu"))
    lat, lon = ds.lat[~ds.lat.isnan()], ds.lon[~ds.lon.isnan()]
    frame = geodesy.HorizontalFrame.from_latlon(lat, lon, ds.scale, ds.offset, 20000.0)
    if variant == "include_height":
        orc = NGPOracle(cfg["pipeline"], None, ds.max_i, fp16=False, geo=(ds.scale, ds.offset, 20000.0))
    else:
        orc = NGPOracle(cfg["pipeline"], frame, ds.max_i, fp16=False)
    params = orc.init_params(3)
    with torch.no_grad():
        for k in ("pos_encoder", "surf_encoder"):
            params[k].mul_(3e3)                                             # non-trivial outputs
    idx = torch.arange(5, len(ds), 211)[:40]
    g = torch.Generator().manual_seed(6)
    pts = torch.rand(64, 3, dtype=torch.float64, generator=g) * 1.6 - 0.8
    job, out = str(tmp_path / "job.pt"), str(tmp_path / "ref_ngp.pt")
    torch.save({"params": {k: v.detach() for k, v in params.items()}, "idx": idx, "seed": 17, "pts": pts}, job)
    r = subprocess.run([sys.executable, "-c", NGP_CHILD, os.path.join(ROOT, "tests", "golden"),
                        os.path.join(ROOT, "atmospheric-neural-rendering_b200", "atmonr", "datasets", "granule.py"),
                        spec, json.dumps(cfg), str(tmp_path), out, ROOT, job], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = torch.load(out, weights_only=False)
    batch = ds[idx]
    torch.manual_seed(17)
    u = torch.rand(idx.shape[0], 24)                                        # samplers.py:37 draws (B, n_bins)
    res = orc.forward(batch, params, u)
    loss = orc.loss(batch, res)
    loss.backward()
    if variant == "include_height":
        assert "norm_heights_fine" in ref["res"]                            # instant_ngp.py:205-206
        ref["res"].pop("norm_heights_fine")
    assert set(ref["res"]) == set(res) - {"pts01"}                          # same result keys (instant_ngp.py:193-204)
    for k, v in ref["res"].items():
        assert res[k].shape == v.shape, k
        assert torch.allclose(res[k].detach(), v, rtol=1e-6, atol=1e-9), (k, float((res[k].detach() - v).abs().max()))
    assert abs(float(loss.detach()) - float(ref["loss"])) <= 1e-6 * abs(float(ref["loss"]))
    for name, gref in ref["grads"].items():
        got = params[name].grad
        assert float((got - gref).abs().max()) <= 1e-5 * float(gref.abs().max() + 1e-30), name
    sig = orc.extract(pts, params)
    assert torch.allclose(sig.detach(), ref["extract"].float(), rtol=1e-6, atol=1e-9)
    # checkpoint layout and optimizer groups of the reference (instant_ngp.py:107-127, 265-296)
    assert ref["state_keys"] == {n: ["params"] for n in ("pos_encoder", "pos_mlp", "dir_encoder", "dir_mlp", "surf_encoder", "surf_mlp")}
    assert [(g[1], g[2]) for g in ref["groups"]] == [(0, 0.01), (0.01, 0.01)]


def test_ngp_training_curve_matches_the_reference_pipeline(tmp_path):
    """30 optimisation steps of the reference's own InstantNGPPipeline + its AdamW groups
    (instant_ngp.py:107-127, trainer.py:99-105) around the stand-in tcnn, against the oracle's
    `train_step` from the same parameters, batches and draws: the two loss curves coincide."""
    from atmonr.datasets.harp2 import HARP2Dataset
    from oracle import geodesy
    from oracle.ngp import NGPOracle
    spec = "synthetic:H=10,W=9,seed=4"
    cfg = json.load(open(os.path.join(ROOT, "configs", "instant_ngp.json")))
    cfg["pipeline"]["num_samples_per_ray"] = 24
    for key in ("encoding", "surface_encoding"):
        cfg["pipeline"]["instant_ngp"][key]["log2_hashmap_size"] = 12
    ds = HARP2Dataset(dict(cfg["dataset"]), spec, device=torch.device("cpu"))
    lat, lon = ds.lat[~ds.lat.isnan()], ds.lon[~ds.lon.isnan()]
    frame = geodesy.HorizontalFrame.from_latlon(lat, lon, ds.scale, ds.offset, 20000.0)
    orc = NGPOracle(cfg["pipeline"], frame, ds.max_i, fp16=False)
    params = orc.init_params(5)
    idx = torch.arange(5, len(ds), 211)[:40]
    steps = 30
    job, out = str(tmp_path / "job.pt"), str(tmp_path / "ref_ngp.pt")
    torch.save({"params": {k: v.detach() for k, v in params.items()}, "idx": idx, "seed": 40,
                "pts": torch.zeros(4, 3, dtype=torch.float64), "steps": steps}, job)
    r = subprocess.run([sys.executable, "-c", NGP_CHILD, os.path.join(ROOT, "tests", "golden"),
                        os.path.join(ROOT, "atmospheric-neural-rendering_b200", "atmonr", "datasets", "granule.py"),
                        spec, json.dumps(cfg), str(tmp_path), out, ROOT, job], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    ref_curve = torch.load(out, weights_only=False)["curve"]
    # the child ran one forward/backward before the loop WITHOUT stepping: parameters are still the initial ones
    opt = orc.make_optimizer(params, cfg["trainer"]["optimizer"])
    curve = []
    for k in range(steps):
        torch.manual_seed(40 + 1 + k)
        u = torch.rand(idx.shape[0], 24)
        loss, _ = orc.train_step(ds[idx + k], params, opt, u)
        curve.append(float(loss))
    assert len(ref_curve) == steps
    rel = max(abs(a - b) / abs(b) for a, b in zip(curve, ref_curve))
    assert rel <= 1e-3, (rel, curve[-3:], ref_curve[-3:])
    assert ref_curve[-1] < ref_curve[0]                     # and it does learn


LOADER_CHILD = DATASET_CHILD.split("idx = torch.arange(0, len(ds), 97)")[0] + r"""
from atmonr.batch_loader import BatchLoader
out = {}
for tag, kw in (("plain", {}), ("drop", {"drop_last": True}), ("seq", {"shuffle": False})):
    torch.manual_seed(123)
    loader = BatchLoader(ds, batch_size=1000, **kw)
    epochs = []
    for _ in range(2):
        epochs.append([b["idx"].clone() for b in loader])
    out[tag] = {"len": len(loader), "epochs": epochs, "after": torch.rand(3)}
torch.save(out, sys.argv[6])
"""


def test_batch_loader_reproduces_the_reference_order(tmp_path):
    """Same `torch.manual_seed`, same batches as the reference's BatchLoader (batch_loader.py:11-60):
    order within two epochs, the partial last batch, drop_last, the sequential mode, len(), and the
    state the global generator is left in."""
    from atmonr.batch_loader import BatchLoader
    from atmonr.datasets.harp2 import HARP2Dataset
    spec = "synthetic:H=10,W=9,seed=4"
    cfg = json.load(open(os.path.join(ROOT, "configs", "instant_ngp.json")))["dataset"]
    out = str(tmp_path / "ref_loader.pt")
    r = subprocess.run([sys.executable, "-c", LOADER_CHILD, os.path.join(ROOT, "tests", "golden"),
                        os.path.join(ROOT, "atmospheric-neural-rendering_b200", "atmonr", "datasets", "granule.py"),
                        spec, json.dumps(cfg), str(tmp_path), out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = torch.load(out, weights_only=False)
    ds = HARP2Dataset(dict(cfg), spec, device=torch.device("cpu"))
    for tag, kw in (("plain", {}), ("drop", {"drop_last": True}), ("seq", {"shuffle": False})):
        torch.manual_seed(123)
        loader = BatchLoader(ds, batch_size=1000, **kw)
        assert len(loader) == ref[tag]["len"], tag
        for want in ref[tag]["epochs"]:
            got = [b["idx"] for b in loader]
            assert len(got) == len(want), tag
            for a, b in zip(got, want):
                assert torch.equal(a.cpu(), b), tag
        assert torch.equal(torch.rand(3), ref[tag]["after"]), tag


NERF_CHILD = r"""
import sys, json, torch
from types import SimpleNamespace
sys.path.insert(0, sys.argv[1])
from make_golden import import_reference
import_reference()
from atmonr.pipelines.nerf import NeRFPipeline
cfg = json.loads(sys.argv[2])
job = torch.load(sys.argv[3])
ds = SimpleNamespace(config={"ray_origin_height": 20000}, scale=job["scale"], offset=job["offset"], max_i=0.3,
                     get_point_preprocessor=lambda name: None)
pipe = NeRFPipeline(cfg, ds)
pipe.load_state_dict(job["params"])
pipe.eval()                                            # no density noise: deterministic given the two draws
torch.manual_seed(job["seed"])
res = pipe.forward(job["batch"])
loss = pipe.compute_loss(job["batch"], res)
loss.backward()
sig = pipe.extract(job["pts"].clone())
torch.save({"res": {k: v.detach() for k, v in res.items()}, "loss": loss.detach(), "extract": sig.detach(),
            "grads": {m: {n: p.grad for n, p in pipe.nerf[m].named_parameters() if p.grad is not None} for m in ("coarse", "fine")}},
           sys.argv[4])
"""


@pytest.mark.parametrize("variant", ["int_L", "include_height"])
def test_nerf_oracle_variants_match_the_reference_pipeline(tmp_path, variant):
    """The NeRF configurations off the shipped config (pipelines/nerf.py:73-217): a scalar `L_x`
    (interleaved encoding layout) without a point preprocessor, and `include_height` (4 encoded
    coordinates): the reference's NeRFPipeline against oracle/nerf.py, forward, loss, gradients, extract."""
    from atmonr.datasets.harp2 import HARP2Dataset
    from oracle import nerf as onerf
    cfgd = json.load(open(os.path.join(ROOT, "configs", "instant_ngp.json")))["dataset"]
    ds = HARP2Dataset(dict(cfgd), "synthetic:H=10,W=9,seed=4", device=torch.device("cpu"))
    cfg = json.load(open(os.path.join(ROOT, "configs", "nerf.json")))["pipeline"]
    cfg.update(mlp_hidden_dim=32, point_preprocessor="", sampler={"N_c": 8, "N_f": 16})
    if variant == "int_L":
        cfg["encoder"] = {"L_x": 5, "L_d": 3}
    else:
        cfg["include_height"], cfg["encoder"] = True, {"L_x": [4, 4, 4, 3], "L_d": 3}
    orc = onerf.NeRFOracle(cfg, None, geo=(ds.scale, ds.offset, 20000.0))
    params = orc.init_params(2)
    idx = torch.arange(3, len(ds), 301)[:20]
    batch = {k: v for k, v in ds[idx].items()}
    g = torch.Generator().manual_seed(1)
    pts = torch.rand(30, 3, dtype=torch.float64, generator=g) * 0.2 + batch["origin"][:1].double()
    job, out = str(tmp_path / "job.pt"), str(tmp_path / "ref.pt")
    torch.save({"params": {m: {k: v.detach() for k, v in params[m].items()} for m in params}, "batch": batch,
                "seed": 31, "pts": pts, "scale": ds.scale, "offset": ds.offset}, job)
    r = subprocess.run([sys.executable, "-c", NERF_CHILD, os.path.join(ROOT, "tests", "golden"), json.dumps(cfg), job, out],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = torch.load(out, weights_only=False)
    torch.manual_seed(31)
    u_c, u_f = torch.rand(20, 8), torch.rand(20, 16)
    res = orc.forward(batch, params, u_c, u_f)
    loss = orc.loss(batch, res)
    loss.backward()
    for k, v in ref["res"].items():
        if k.startswith("norm_heights"):
            continue
        assert torch.allclose(res[k].detach(), v, rtol=2e-5, atol=1e-7), (k, float((res[k].detach() - v).abs().max()))
    assert abs(float(loss.detach()) - float(ref["loss"])) <= 1e-5 * abs(float(ref["loss"]))
    for m in ("coarse", "fine"):
        for n in ("fc1.weight", "fc6.weight", "fc11.weight"):
            a, b = params[m][n].grad, ref["grads"][m][n]
            assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max() + 1e-30), (m, n)
    assert torch.allclose(orc.extract(pts, params).detach(), ref["extract"], rtol=2e-5, atol=1e-7)


# ------------------------------------------------------------------------------------------
# the reference's Trainer, end to end, against this package's Trainer
# ------------------------------------------------------------------------------------------
TRAINER_CHILD = NGP_CHILD.split("from atmonr.datasets.harp2 import HARP2Dataset\nfrom atmonr.pipelines.instant_ngp import InstantNGPPipeline")[0] + r"""
import numpy as np
torch.cuda.current_device = lambda: 0
from atmonr.datasets import harp2 as ref_harp2
from atmonr.datasets.harp2 import HARP2Dataset
from atmonr.pipelines.instant_ngp import InstantNGPPipeline
from atmonr import trainer as ref_trainer

def _psnr(pred, target, dim, reduction, data_range):      # torchmetrics' definition, for the stub
    return 10 * torch.log10(data_range ** 2 / ((pred - target) ** 2).mean(dim=dim))
ref_harp2.peak_signal_noise_ratio = _psnr
ref_harp2.structural_similarity_index_measure = lambda p, t, reduction: torch.zeros(p.shape[0])

class Recorder:
    def __init__(self, *a, **k): self.scalars, self.images = [], []
    def add_scalar(self, tag, val, step): self.scalars.append((tag, float(val), int(step)))
    def add_image(self, tag, img): self.images.append((tag, np.array(img)))
ref_trainer.SummaryWriter = Recorder

cfg = json.loads(sys.argv[4])
ds = HARP2Dataset(cfg["dataset"], "fake.nc")
pipe = InstantNGPPipeline(cfg["pipeline"], ds)
pipe.send_tensors_to("cpu")
job = torch.load(sys.argv[8])
for name, p in job["params"].items():
    getattr(pipe, name).params.data.copy_(p)
torch.manual_seed(job["seed"])
tr = ref_trainer.Trainer(cfg["trainer"], ds, pipe, "t")
outdir = Path(sys.argv[5]) / "ref_run"
outdir.mkdir()
tr.train(outdir)
ck = sorted(outdir.glob("epoch_*.pt"))
last = torch.load(ck[-1], weights_only=False)
torch.save({"scalars": tr.writer.scalars, "images": tr.writer.images, "lr": tr.optimizer.param_groups[0]["lr"],
            "iter_count": tr.iter_count, "epoch_idx": tr.epoch_idx, "num_epochs": tr.num_epochs,
            "ckpts": [c.name for c in ck], "ck_keys": sorted(last), "ck_pipeline": last["pipeline"],
            "params": {n: getattr(pipe, n).params.detach().clone() for n in job["params"]}}, sys.argv[6])
"""


class _OracleBackedPipeline:
    """Test-only adapter: the Pipeline surface this package's Trainer talks to, computed by the oracle
    (the native pipeline needs a GPU; the Trainer's own logic does not)."""

    def __init__(self, orc, params):
        self.orc, self.params, self.device = orc, params, "cpu"

    def get_optimizer(self, cfg):
        return self.orc.make_optimizer(self.params, cfg)

    def forward(self, batch):
        u = torch.rand((batch["origin"].shape[0], self.orc.cfg["num_samples_per_ray"]))   # samplers.py:37
        return self.orc.forward(batch, self.params, u)

    def compute_loss(self, batch, results):
        return self.orc.loss(batch, results)

    def state_dict(self):
        return {k: {"params": v.detach().clone()} for k, v in self.params.items()}


def test_trainer_matches_the_reference_trainer(tmp_path, monkeypatch):
    """trainer.py:26-274 end to end: the reference's Trainer drives its InstantNGPPipeline (stand-in
    tcnn) for two epochs; this package's Trainer drives the oracle from the same seed. Same batches in
    the same order, same per-iteration losses under the same step indices, same learning-rate decays,
    same PSNR and progress image at each epoch end, same checkpoints."""
    from atmonr import trainer as T
    from atmonr.datasets.harp2 import HARP2Dataset
    from oracle import geodesy
    from oracle.ngp import NGPOracle
    spec = "synthetic:H=12,W=12,seed=2"
    cfg = json.load(open(os.path.join(ROOT, "configs", "instant_ngp.json")))
    cfg["pipeline"]["num_samples_per_ray"] = 8
    for key in ("encoding", "surface_encoding"):
        cfg["pipeline"]["instant_ngp"][key]["log2_hashmap_size"] = 10
    cfg["trainer"].update(batch_size=4096, num_iters=7, print_frequency=3,
                          scheduler={"type": "fixed", "gamma": 0.5, "decay_start": 2, "decay_interval": 2})
    ds = HARP2Dataset(dict(cfg["dataset"]), spec, device=torch.device("cpu"))
    lat, lon = ds.lat[~ds.lat.isnan()], ds.lon[~ds.lon.isnan()]
    frame = geodesy.HorizontalFrame.from_latlon(lat, lon, ds.scale, ds.offset, 20000.0)
    orc = NGPOracle(cfg["pipeline"], frame, ds.max_i, fp16=False)
    params = orc.init_params(7)
    with torch.no_grad():
        for k in ("pos_encoder", "surf_encoder"):
            params[k].mul_(3e3)
    job, out = str(tmp_path / "job.pt"), str(tmp_path / "ref_trainer.pt")
    torch.save({"params": {k: v.detach() for k, v in params.items()}, "seed": 99}, job)
    r = subprocess.run([sys.executable, "-c", TRAINER_CHILD, os.path.join(ROOT, "tests", "golden"),
                        os.path.join(ROOT, "atmospheric-neural-rendering_b200", "atmonr", "datasets", "granule.py"),
                        spec, json.dumps(cfg), str(tmp_path), out, ROOT, job], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = torch.load(out, weights_only=False)

    class Recorder:
        def __init__(self):
            self.scalars, self.images = [], []

        def add_scalar(self, tag, val, step):
            self.scalars.append((tag, float(val), int(step)))

        def add_image(self, tag, img):
            self.images.append((tag, img))

    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(T, "_make_writer", lambda d: Recorder())
    pipe = _OracleBackedPipeline(orc, params)
    torch.manual_seed(99)
    tr = T.Trainer(cfg["trainer"], ds, pipe, "t")
    outdir = tmp_path / "my_run"
    outdir.mkdir()
    tr.train(outdir)
    assert (tr.iter_count, tr.epoch_idx, tr.num_epochs) == (ref["iter_count"], ref["epoch_idx"], ref["num_epochs"])
    assert tr.optimizer.param_groups[0]["lr"] == pytest.approx(ref["lr"])
    mine_loss = [(s, v) for t, v, s in tr.writer.scalars if t == "Loss"]
    ref_loss = [(s, v) for t, v, s in ref["scalars"] if t == "Loss"]
    assert [s for s, _ in mine_loss] == [s for s, _ in ref_loss] == list(range(7))
    for (_, a), (_, b) in zip(mine_loss, ref_loss):
        assert abs(a - b) <= 2e-4 * abs(b), (mine_loss, ref_loss)
    mine_psnr = [v for t, v, s in tr.writer.scalars if t == "PSNR_mean"]
    ref_psnr = [v for t, v, s in ref["scalars"] if t == "PSNR_mean"]
    assert len(mine_psnr) == len(ref_psnr) == 2
    assert all(abs(a - b) <= 1e-3 * abs(b) for a, b in zip(mine_psnr, ref_psnr))
    assert [t for t, _ in tr.writer.images] == [t for t, _ in ref["images"]]
    for (_, a), (_, b) in zip(tr.writer.images, ref["images"]):
        assert a.shape == b.shape and abs(a - b).max() <= 1e-4
    ck = sorted(outdir.glob("epoch_*.pt"))
    assert [c.name for c in ck] == ref["ckpts"]
    last = torch.load(ck[-1], weights_only=False)
    assert sorted(last) == ref["ck_keys"]
    for name, p in ref["params"].items():
        assert params[name].shape == p.shape, name
        if p.numel():                                       # dir_encoder has no parameters
            assert float((params[name].detach() - p).abs().max()) <= 2e-4 * float(p.abs().max() + 1e-30), name
        assert list(last["pipeline"][name]) == list(ref["ck_pipeline"][name]) == ["params"]


GRID_CHILD = DATASET_CHILD.split("idx = torch.arange(0, len(ds), 97)")[0] + r"""
from atmonr.datasets import harp2_extract as ref_ext
ref_ext.HARP2VoxelGridExtractDataset._interp_dem_height = lambda self, dem, la, lo: torch.zeros_like(la)   # no DEM file
kw = json.loads(sys.argv[7])
grid = ref_ext.HARP2VoxelGridExtractDataset(ds, **kw)
torch.save({"shp": tuple(grid.shp), "lat": grid.lat, "lon": grid.lon, "xyz": grid.xyz, "idx": grid.idx,
            "sample_alt": grid.sample_alt, "batch": grid.__getbatch__(torch.arange(0, 50, 7))}, sys.argv[6])
"""


@pytest.mark.parametrize("spec,kw", [
    ("synthetic:H=10,W=9,seed=4", {"horizontal_step": 30000.0, "alt_step": 1000.0}),
    ("synthetic:H=10,W=9,seed=4,lat0=-62,lon0=177.5", {"horizontal_step": 21000.0, "alt_step": 2500.0, "min_alt": 500.0, "max_alt": 15000.0}),
])
def test_vincenty_voxel_grid_matches_the_reference_layout(tmp_path, spec, kw):
    """HARP2VoxelGridExtractDataset(layout="vincenty") against the reference class
    (harp2_extract.py:189-348; its DEM lookup, which only feeds the output file's `height`, stubbed):
    same grid shape and altitude levels, voxel columns within a metre (the geodesic arithmetic is
    float32 on both sides, in different operation orders), same index table and batch layout."""
    from atmonr.datasets.harp2 import HARP2Dataset
    from atmonr.datasets.harp2_extract import HARP2VoxelGridExtractDataset
    cfg = json.load(open(os.path.join(ROOT, "configs", "instant_ngp.json")))["dataset"]
    out = str(tmp_path / "ref_grid.pt")
    r = subprocess.run([sys.executable, "-c", GRID_CHILD, os.path.join(ROOT, "tests", "golden"),
                        os.path.join(ROOT, "atmospheric-neural-rendering_b200", "atmonr", "datasets", "granule.py"),
                        spec, json.dumps(cfg), str(tmp_path), out, json.dumps(kw)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = torch.load(out, weights_only=False)
    ds = HARP2Dataset(dict(cfg), spec, device=torch.device("cpu"))
    grid = HARP2VoxelGridExtractDataset(ds, layout="vincenty", **kw)
    assert tuple(grid.shp) == ref["shp"] and torch.equal(grid.sample_alt, ref["sample_alt"])
    assert torch.equal(grid.idx, ref["idx"]) and grid.xyz.dtype == ref["xyz"].dtype == torch.float64
    dlon = (grid.lon - ref["lon"] + 180) % 360 - 180
    assert float((grid.lat - ref["lat"]).abs().max()) <= 2e-5 and float(dlon.abs().max()) <= 2e-5     # ~2 m
    assert float((grid.xyz - ref["xyz"]).norm(dim=1).max()) <= 3.0                                     # metres
    b = grid.__getbatch__(torch.arange(0, 50, 7))
    assert set(b) == set(ref["batch"]) and torch.equal(b["idx"], ref["batch"]["idx"])
    # neighbouring columns are about `horizontal_step` apart: n stations span n steps (the reference's
    # linspace), and the meridians converge towards the poleward edge (16 % over 5 degrees at 62 S)
    xyz = grid.xyz.view(*grid.shp, 3)[:, :, 0]
    step_ew = (xyz[:, 1:] - xyz[:, :-1]).norm(dim=-1)
    assert float((step_ew / kw["horizontal_step"] - 1).abs().max()) < 0.25


# ------------------------------------------------------------------------------------------
# the other three coordinate modes of scripts/extract.py: L1C, EarthCARE, global grid
# ------------------------------------------------------------------------------------------
_L1B_NAME = "PACE_HARP2.20240101T000000.L1B.V3.nc"          # the L1C class parses this (harp2_extract.py:141)
LAYOUT_CHILD = (DATASET_CHILD.split("idx = torch.arange(0, len(ds), 97)")[0]
                .replace("fake.nc", _L1B_NAME)
                .replace("    def __init__(self, path): pass\n", "    def __init__(self, path): self.path = str(path)\n")
                .replace("        group, name = path.split(\"/\")\n",
                         "        group, name = path.split(\"/\")\n"
                         "        if \"L1C\" in self.path: return Var(gran.l1c_geolocation()[name])\n")) + r"""
import h5py
from atmonr.datasets import harp2_extract as ref_ext
mode, kw = sys.argv[7], json.loads(sys.argv[8])
out = {}
if mode == "l1c":
    Path("data/HARP2_L1C").mkdir(parents=True)
    Path("data/HARP2_L1C/PACE_HARP2.20240101T000000.L1C.V3.5km.nc").touch()      # present: no download
    grid = ref_ext.HARP2L1CExtractDataset(ds, **kw)
    out = {"shp": tuple(grid.shp), "lat": grid.lat, "lon": grid.lon, "height": grid.height, "sample_alt": grid.sample_alt}
elif mode == "earthcare":
    track = gran.earthcare_track()
    class H5Var:
        def __init__(self, a): self.a = a
        def __getitem__(self, k): return self.a[k] if not isinstance(self.a, bytes) else self.a
    h5py.Dataset = H5Var
    table = {"HeaderData/FixedProductHeader/File_Type": H5Var(kw.pop("file_type", track["file_type"]).encode()),
             "ScienceData/height": H5Var(track["height"]), "ScienceData/latitude": H5Var(track["latitude"]),
             "ScienceData/longitude": H5Var(track["longitude"])}
    h5py.File = lambda path: table
    try:
        grid = ref_ext.HARP2EarthCAREExtractDataset(ds, **kw)
        out = {"shp": tuple(grid.shp), "lat": grid.lat, "lon": grid.lon, "alt": grid.alt}
    except NotImplementedError as err:
        torch.save({"error": str(err)}, sys.argv[6]); sys.exit(0)
else:
    # the reference's constructor cannot finish (harp2_extract.py:896 chains two tensor comparisons):
    # stop it at the altitude computation that precedes that line and keep what it has built
    class Reached(Exception): pass
    def stop(x, y, z): raise Reached
    ref_ext.cartesian_to_horizontal = stop
    ref_ext.tqdm = lambda it, **k: it
    grid = object.__new__(ref_ext.HARP2GlobalGridExtractDataset)
    try:
        ref_ext.HARP2GlobalGridExtractDataset.__init__(grid, ds, **kw)
        raise SystemExit("the reference's global grid finished: update this test")
    except Reached:
        pass
    out = {"voxels": grid.voxels}
    grid.idx = torch.arange(grid.xyz.shape[0], dtype=torch.int32)
out.update(xyz=grid.xyz, idx=grid.idx, batch=grid.__getbatch__(torch.arange(0, 40, 7)))
torch.save(out, sys.argv[6])
"""


def _reference_layout(tmp_path, spec, mode, kw):
    cfg = json.load(open(os.path.join(ROOT, "configs", "instant_ngp.json")))["dataset"]
    out = str(tmp_path / "ref_layout.pt")
    r = subprocess.run([sys.executable, "-c", LAYOUT_CHILD, os.path.join(ROOT, "tests", "golden"),
                        os.path.join(ROOT, "atmospheric-neural-rendering_b200", "atmonr", "datasets", "granule.py"),
                        spec, json.dumps(cfg), str(tmp_path), out, mode, json.dumps(kw)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    from atmonr.datasets.harp2 import HARP2Dataset
    return torch.load(out, weights_only=False), HARP2Dataset(dict(cfg), spec, device=torch.device("cpu"))


@pytest.mark.parametrize("spec,kw", [
    ("synthetic:H=10,W=9,seed=4", {"alt_step": 2500.0}),
    ("synthetic:H=10,W=9,seed=4,lat0=-62,lon0=177.5", {"alt_step": 1000.0, "min_alt": 500.0, "max_alt": 9000.0}),
])
def test_l1c_layout_matches_the_reference_class(tmp_path, spec, kw):
    """HARP2L1CExtractDataset (harp2_extract.py:115-186) fed the same L1C geolocation through a stand-in
    netCDF4.Dataset: the point table, the index table and the per-bin fields bit for bit (fill values as
    NaN, north first)."""
    from atmonr.datasets.factory import get_extract_dataset
    ref, ds = _reference_layout(tmp_path, spec, "l1c", kw)
    grid = get_extract_dataset("L1C", ds, horizontal_step=3000.0, scale=1.0, **kw)     # the script passes every flag
    assert tuple(grid.shp) == ref["shp"] and torch.equal(grid.sample_alt, ref["sample_alt"])
    assert grid.xyz.dtype == torch.float64 and grid.xyz.shape == ref["xyz"].shape
    for name in ("lat", "lon", "height", "xyz"):
        got, want = getattr(grid, name), ref[name]
        assert got.dtype == want.dtype and torch.equal(got.isnan(), want.isnan()), name
        assert torch.equal(got.nan_to_num(), want.nan_to_num()), name
    assert bool(grid.xyz.isnan().any()) and float(grid.lat[0, 5, 0]) > float(grid.lat[-1, 5, 0])   # fill bins; north first
    assert torch.equal(grid.idx, ref["idx"])
    b = grid.__getbatch__(torch.arange(0, 40, 7))
    assert set(b) == set(ref["batch"]) and torch.equal(b["idx"], ref["batch"]["idx"])
    # the voxel grid's writer serves this layout as well
    sigma = torch.rand(grid.xyz.shape[0], 1)
    grid.dump(tmp_path / "l1c_extract.nc", sigma)
    stored = np.load(tmp_path / "l1c_extract.npz")
    assert stored["extinction_coefficient"].shape == (*ref["shp"], grid.sample_alt.shape[0], 1)
    assert np.array_equal(stored["height"], ref["height"].numpy(), equal_nan=True) and tuple(grid.shp) == ref["shp"]


@pytest.mark.parametrize("spec,kw", [
    ("synthetic:H=10,W=9,seed=4", {"earthcare_filename": "synthetic", "earthcare_range": None}),
    ("synthetic:H=10,W=9,seed=4,lat0=-62,lon0=177.5", {"earthcare_filename": "synthetic", "earthcare_range": [12, 131]}),
])
def test_earthcare_layout_matches_the_reference_class(tmp_path, spec, kw):
    """HARP2EarthCAREExtractDataset (harp2_extract.py:599-675) fed the same ATLID track through a stand-in
    h5py.File: profile range, the all-profiles altitude mask, the curtain's point table: bit for bit."""
    from atmonr.datasets.factory import get_extract_dataset
    ref, ds = _reference_layout(tmp_path, spec, "earthcare", kw)
    grid = get_extract_dataset("earthcare", ds, alt_step=250.0, **kw)
    assert tuple(grid.shp) == ref["shp"] and grid.shp[1] < 90          # range bins outside the shell were dropped
    for name in ("lat", "lon", "alt"):
        assert np.array_equal(getattr(grid, name), ref[name]), name
    assert grid.xyz.dtype == ref["xyz"].dtype == torch.float64 and torch.equal(grid.xyz, ref["xyz"])
    assert torch.equal(grid.idx, ref["idx"])
    b = grid.__getbatch__(torch.arange(0, 40, 7))
    assert torch.equal(b["xyz"], ref["batch"]["xyz"]) and torch.equal(b["idx"], ref["batch"]["idx"])
    grid.dump(tmp_path / "curtain.nc", torch.rand(grid.xyz.shape[0], 4))
    stored = np.load(tmp_path / "curtain.npz")
    assert stored["extinction_coefficient"].shape == (*ref["shp"], 4) and stored["height"].shape == ref["shp"]
    assert np.array_equal(stored["longitude"], ref["lon"][:, 0]) and float(stored["attr_neural_rendering_scene_scale"]) == ds.scale
    with pytest.raises(AssertionError):
        get_extract_dataset("earthcare", ds, earthcare_filename="synthetic", earthcare_range=[5, 5])


def test_earthcare_layout_rejects_other_products_like_the_reference(tmp_path, monkeypatch):
    from atmonr.datasets import granule
    from atmonr.datasets.factory import get_extract_dataset
    spec = "synthetic:H=10,W=9,seed=4"
    ref, ds = _reference_layout(tmp_path, spec, "earthcare", {"earthcare_filename": "synthetic", "earthcare_range": None, "file_type": "ATL_NOM_1B"})
    track = granule.SyntheticGranule.earthcare_track
    monkeypatch.setattr(granule.SyntheticGranule, "earthcare_track", lambda self: {**track(self), "file_type": "ATL_NOM_1B"})
    with pytest.raises(NotImplementedError) as err:
        get_extract_dataset("earthcare", ds, earthcare_filename="synthetic", earthcare_range=None)
    assert str(err.value) == ref["error"]
    # a product file: the .npz stand-in for the HDF5 file, same variable paths
    monkeypatch.chdir(tmp_path)
    os.makedirs("data/EarthCARE", exist_ok=True)
    t = track(ds.granule)
    np.savez(os.path.join("data", "EarthCARE", "ECA_EXAE_ATL_EBD_2A_x.npz"), **{
        "HeaderData/FixedProductHeader/File_Type": np.bytes_(b"ATL_EBD_2A"), "ScienceData/height": t["height"],
        "ScienceData/latitude": t["latitude"], "ScienceData/longitude": t["longitude"]})
    monkeypatch.setattr(granule.SyntheticGranule, "earthcare_track", track)
    from_file = get_extract_dataset("earthcare", ds, earthcare_filename="ECA_EXAE_ATL_EBD_2A_x.h5", earthcare_range=None)
    synthetic = get_extract_dataset("earthcare", ds, earthcare_filename="synthetic", earthcare_range=None)
    assert torch.equal(from_file.xyz, synthetic.xyz)


@pytest.mark.parametrize("spec,kw", [
    ("synthetic:H=10,W=9,seed=4", {"scale": 100 / 6.378e6, "grid_res": 0.4, "vstretch": 12, "lon_crop": 0.05}),
    ("synthetic:H=8,W=7,seed=2,lat0=-42,lon0=100", {"scale": 100 / 6.378e6, "grid_res": 0.25, "vstretch": None, "lon_crop": 0.1}),
])
def test_global_grid_layout_matches_the_reference_class(tmp_path, spec, kw):
    """HARP2GlobalGridExtractDataset (harp2_extract.py:794-903) up to the line the reference cannot
    execute (:896): the voxels crossed by the granule's rays in the stretched spherical frame, the
    per-layer longitude crop, the un-stretched WGS-84 voxel centres: the same SET of (voxel, centre)
    rows. This package then applies the cull the reference's comment describes."""
    from oracle import geodesy
    from atmonr.datasets.factory import get_extract_dataset
    ref, ds = _reference_layout(tmp_path, spec, "globalgrid", kw)
    grid = get_extract_dataset("globalgrid", ds, alt_step=250.0, **kw)
    assert ref["voxels"].dtype == grid.voxels.dtype == torch.int32 and ref["xyz"].dtype == grid.xyz.dtype == torch.float32
    alt = geodesy.ecef_to_geodetic(*(ref["xyz"][:, k] for k in range(3)))[2]
    keep = ~((alt <= 0) | (alt > 20000.0))
    assert 0 < int(keep.sum()) <= keep.numel() and ref["voxels"].shape[0] > 50
    rows = lambda vox, xyz: sorted(map(tuple, torch.cat([vox.double(), xyz.double()], dim=1).tolist()))
    assert rows(grid.voxels, grid.xyz) == rows(ref["voxels"][keep], ref["xyz"][keep])
    assert torch.equal(grid.idx, torch.arange(int(keep.sum()), dtype=torch.int32))
    # every kept voxel centre lies inside the shell and (un-cropped) under the granule's rays
    got_alt = geodesy.ecef_to_geodetic(*(grid.xyz.double()[:, k] for k in range(3)))[2]
    assert float(got_alt.min()) > 0 and float(got_alt.max()) <= 20000.0
    # the reference's own fallback writer (no OpenVDB bindings): voxels.npy + sigma.npy
    sigma = torch.rand(grid.xyz.shape[0], 1)
    grid.dump(tmp_path / "grid.vdb", sigma)
    assert np.array_equal(np.load(tmp_path / "voxels.npy"), grid.voxels.numpy())
    assert np.array_equal(np.load(tmp_path / "sigma.npy"), sigma.numpy())
    with pytest.raises(FileExistsError):
        grid.dump(tmp_path / "grid.vdb", sigma)
    with pytest.raises(AssertionError):
        get_extract_dataset("globalgrid", ds, scale=1.0, grid_res=1.0, vstretch=0.5)


def test_extract_modes_are_the_reference_registry():
    """datasets/factory.py:24-33: the four coordinate modes, case-insensitive; anything else raises."""
    from atmonr.datasets import factory
    assert sorted(factory._EXTRACT_DATASETS["HARP2"]) == ["earthcare", "globalgrid", "l1c", "voxelgrid"]
    ds = FakeDataset(tiny_scene())
    ds.config = {"type": "HARP2", "ray_origin_height": 20000}
    with pytest.raises(NotImplementedError):
        factory.get_extract_dataset("octree", ds)
    ds.config = {"type": "AirHARP"}
    with pytest.raises(NotImplementedError):
        factory.get_extract_dataset("voxelgrid", ds)


REF_SCRIPTS = "/root/reference/scripts"


@pytest.mark.parametrize("script", ["train.py", "extract.py"])
def test_reference_scripts_run_unchanged_against_this_package(script):
    """north_star: "scripts/train.py and scripts/extract.py run unchanged". The REFERENCE's own script files
    (read where they lie, never copied) are executed with this package as the only `atmonr` on the path:
    every import they make (atmonr.datasets.factory: BANDS / get_dataset / get_extract_dataset,
    atmonr.pipelines.factory, atmonr.trainer.Trainer, atmonr.batch_loader, atmonr.utils.load_config,
    atmonr.geospatial.spherical.EARTH_RADIUS) must resolve here, and their argument parser must come up.
    The same command-line surface is offered by this repo's scripts/ (flags compared below). Running the
    reference scripts further needs a GPU, which the box that has one does not have the reference tree for;
    tests/test_gpu_e2e.py drives this repo's scripts (same flags, same call sequence) on the GPU."""
    import re
    pkg = os.path.join(ROOT, "atmospheric-neural-rendering_b200")
    env = dict(os.environ, PYTHONPATH=pkg)
    r = subprocess.run([sys.executable, os.path.join(REF_SCRIPTS, script), "--help"], capture_output=True, text=True,
                       timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    ref_flags = set(re.findall(r"--[a-z][a-z-]+", r.stdout))
    assert "--exp-name" in ref_flags
    mine = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script), "--help"], capture_output=True,
                          text=True, timeout=300, env=dict(os.environ, PYTHONPATH=""), cwd=ROOT)
    assert mine.returncode == 0, mine.stderr[-2000:]
    my_flags = set(re.findall(r"--[a-z][a-z-]+", mine.stdout))
    assert ref_flags <= my_flags, ref_flags - my_flags          # every reference flag exists here too
    # which atmonr did the reference script import?
    probe = ("import sys, runpy; sys.argv=['x','--help']\n"
             "try:\n    runpy.run_path(%r, run_name='__main__')\nexcept SystemExit:\n    pass\n"
             "import atmonr; print('ATMONR_FROM', atmonr.__file__)") % os.path.join(REF_SCRIPTS, script)
    r = subprocess.run([sys.executable, "-c", probe], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    assert f"ATMONR_FROM {pkg}" in r.stdout
