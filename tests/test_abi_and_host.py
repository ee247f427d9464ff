"""CPU-side checks: the shared library exports every symbol the header declares, host-only entry
points work without a GPU, and the exact per-sample code the kernels run (built for the host in
libatmonr_hostcheck.so) agrees with the oracle."""

import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from helpers import ROOT, ngp_config, take, tiny_scene
from oracle import geodesy, sampling, tcnn_spec
from oracle.ngp import NGPOracle


@pytest.fixture(scope="module", autouse=True)
def built():
    import __graft_entry__ as ge
    ge.build()


def _header_functions():
    text = open(os.path.join(ROOT, "include", "atmonr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(atmonr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from atmonr.native import lib as L
    lib = L.load()
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/atmonr_b200.h but not exported"
    assert sorted(L.SIGNATURES) == names, "ctypes SIGNATURES out of sync with the header"
    assert lib.atmonr_abi_version() == 1


def test_no_cpu_fallback():
    from atmonr.native import lib as L, ops
    with pytest.raises(L.NativeLibraryError):
        ops.sample_uniform(torch.zeros(2, 3), torch.zeros(2, 3), torch.ones(2), 4, random=False)
    from atmonr import losses
    with pytest.raises(L.NativeLibraryError):
        losses.mse_loss(torch.rand(4), torch.rand(4), 1.0)


def test_bad_arguments_report_errors():
    from atmonr.native import lib as L
    g = L.GridT()
    with pytest.raises(L.NativeLibraryError, match="n_dims"):
        L.call("atmonr_grid_layout", 5, 16, 19, 16, 1.5, C.byref(g))
    with pytest.raises(L.NativeLibraryError, match="n_levels"):
        L.call("atmonr_grid_layout", 3, 17, 19, 16, 1.5, C.byref(g))
    # argument checks of the dataset-side entry points happen before any CUDA call
    with pytest.raises(L.NativeLibraryError, match="atmonr_get_rays: null pointer"):
        L.call("atmonr_get_rays", None, None, None, None, None, 4, 20000.0, 10.0, 20, None, None, None, None, None, None)
    with pytest.raises(L.NativeLibraryError, match="atmonr_gather_batch"):
        L.call("atmonr_gather_batch", *([None] * 8), 4, 0, *([None] * 9))
    L.call("atmonr_get_rays", None, None, None, None, None, 0, 20000.0, 10.0, 20, None, None, None, None, None, None)   # empty chunk
    L.call("atmonr_gather_batch", *([None] * 8), 0, 10, *([None] * 9))                                                 # empty batch


def test_native_ray_dispatch_needs_cuda_tensors(monkeypatch):
    """The ray-setup kernels only take CUDA inputs; CPU tensors keep the torch expressions (dataset
    construction without a GPU, e.g. this test-suite), and the operator itself refuses CPU tensors."""
    from atmonr.geospatial.wgs_84 import get_rays
    from atmonr.native import lib as L, ops
    t = lambda v: torch.tensor([[v]], dtype=torch.float32)
    args = (t(35.0), t(-75.0), t(0.0), t(10.0), t(0.0))
    want = get_rays(*args, 20000.0)
    monkeypatch.setenv("ATMONR_NATIVE_RAYS", "1")
    got = get_rays(*args, 20000.0)
    assert all(torch.equal(a, b) for a, b in zip(want, got))
    with pytest.raises(L.NativeLibraryError):
        ops.get_rays(*args, 20000.0)
    with pytest.raises(L.NativeLibraryError):
        ops.gather_batch({k: torch.zeros(2, 3) if k in ("origin", "dir") else torch.zeros(2, dtype=d)
                          for k, d in (("origin", None), ("dir", None), ("alt", torch.float32), ("rad", torch.float32),
                                       ("len", torch.float32), ("idx", torch.int32), ("irgb_idx", torch.int64))},
                         torch.tensor([0]))


@pytest.mark.parametrize("dims,key", [(3, "encoding"), (2, "surface_encoding"), (4, "encoding")])
def test_grid_layout_matches_oracle_bit_for_bit(dims, key):
    from atmonr.native import lib as L
    cfg = ngp_config()["instant_ngp"][key]
    if dims == 2:
        cfg = cfg["nested"][0]
    g = L.grid_layout(dims, cfg)
    lv = tcnn_spec.grid_levels(dims, 16, cfg["log2_hashmap_size"], cfg["base_resolution"], cfg["per_level_scale"])
    assert np.array_equal(np.array(g.scale[:16], dtype=np.float32).view(np.uint32), lv["scale"].view(np.uint32))
    assert list(g.res[:16]) == lv["res"].tolist() and list(g.size[:16]) == lv["size"].tolist()
    assert list(g.offset[:17]) == lv["offset"].tolist()
    if dims < 4:  # 4 = positions + height (`include_height`)
        assert g.n_entries == (21141696 if dims == 3 else 2761000)


def test_level_table_does_not_depend_on_how_tcnn_rounds_the_level_scale():
    """Round-1 review: oracle/tcnn_spec.py evaluates `exp2(l * log2(per_level_scale)) * base - 1` in float64
    and rounds once, tiny-cuda-nn evaluates it in float32 (`exp2f`, `log2f`; on the device `exp2f` is a 2-ulp
    approximation), so the two scales can differ in the last bits. Quantified here for the shipped grids
    (configs/instant_ngp.json: 3-D table 2^21, 2-D table 2^19, and the 4-D `include_height` table):
    * the literal float32 evaluation is within 8 ulps of the once-rounded scale at every level;
    * no level's scale lies within 500 ulps of an integer (level 0 is exactly 15 under any evaluation), so
      `res = ceil(scale) + 1`, and with it every level size, every offset and which levels hash, is the
      same for ANY evaluation that close: the TABLE LAYOUT cannot differ from tiny-cuda-nn's;
    * what can differ is `pos = scale * x + 0.5` by <= 8 ulps of the scale: 1e-3 of a cell at the finest level
      (scale 2047), i.e. an interpolation-weight change below the fp16 resolution of the features."""
    import json
    from oracle import tcnn_spec
    cfg = json.load(open(os.path.join(ROOT, "configs", "instant_ngp.json")))["pipeline"]["instant_ngp"]
    grids = [(3, cfg["encoding"]), (4, cfg["encoding"]), (2, cfg["surface_encoding"]["nested"][0])]
    f32 = np.float32
    for dims, g in grids:
        assert g["otype"] == "HashGrid"
        lv = tcnn_spec.grid_levels(dims, g["n_levels"], g["log2_hashmap_size"], g["base_resolution"], g["per_level_scale"])
        pls, base = f32(g["per_level_scale"]), f32(g["base_resolution"])
        literal = np.array([f32(np.exp2(f32(l) * np.log2(pls))) * base - f32(1) for l in range(g["n_levels"])], dtype=f32)
        ulp = np.spacing(lv["scale"]).astype(np.float64)
        off_by = np.abs(literal.astype(np.float64) - lv["scale"].astype(np.float64)) / ulp
        assert off_by.max() <= 8
        to_integer = np.abs(lv["scale"].astype(np.float64) - np.round(lv["scale"].astype(np.float64))) / ulp
        assert lv["scale"][0] == 15.0 and literal[0] == 15.0 and to_integer[1:].min() > 500
        assert np.array_equal(np.ceil(literal).astype(np.uint32) + 1, lv["res"])
        assert float(off_by.max() * ulp[-1]) <= 1.1e-3          # cells, at x = 1 on the finest level


def _hc():
    lib = C.CDLL(os.path.join(ROOT, "atmospheric-neural-rendering_b200", "lib", "libatmonr_hostcheck.so"))
    return lib


def _frame_struct(fr):
    from atmonr.native import lib as L
    return L.make_frame(fr.scale, fr.offset, fr.lat_min, fr.lat_range, fr.lon_min, fr.lon_range, fr.origin_height, fr.shift_lon)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_host_build_of_kernel_math_matches_oracle():
    scene = tiny_scene()
    b = take(scene.batch, slice(0, 64))
    n = 48
    u = torch.rand(64, n, generator=torch.Generator().manual_seed(0))
    orc = NGPOracle(ngp_config(n), scene.frame, scene.max_i)
    want = orc.forward(b, orc.init_params(0), u)
    hc, fr = _hc(), _frame_struct(scene.frame)
    x01 = np.zeros((64 * n, 3), np.float32)
    z = np.zeros((64, n), np.float32)
    bins = torch.linspace(0, 1, n + 1)[:-1].numpy().copy()
    o, d, ln, un = (t.numpy().copy() for t in (b["origin"], b["dir"], b["len"], u))
    hc.hc_ngp_sample_points(C.byref(fr), _p(o), _p(d), _p(ln), _p(un), _p(bins), C.c_int64(64), n, 1,
                            C.c_uint64(0), C.c_uint64(0), C.c_float(8.0), _p(x01), _p(z))
    assert np.array_equal(z, want["z_vals_fine"].numpy())
    assert np.abs(x01 - want["pts01"].numpy()).max() < 1.5e-7
    # float64 flavour
    p64 = (torch.rand(500, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(1)) * 2 - 1).numpy().copy()
    out64 = np.zeros_like(p64)
    hc.hc_preprocess_f64(C.byref(fr), _p(p64), _p(out64), C.c_int64(500))
    want64 = geodesy.preprocess_horizontal(torch.from_numpy(p64)[None], scene.frame)[0].numpy()
    assert np.abs(out64 - want64).max() < 1e-11


@pytest.mark.parametrize("lat0,lon0,dlat,dlon,shift", [
    (32.5, -72.5, 5, 5, False),     # the bench granule
    (-60.0, 179.5, 4, 6, True),     # across the dateline (harp2.py:366-370)
    (75.0, 100.0, 8, 19, False),    # wide swath at high latitude: partly outside the small-angle window
    (10.0, -120.0, 30, 40, False),  # mostly outside it: literal atan2 form
])
def test_local_angle_geodesy_matches_literal_form(lat0, lon0, dlat, dlon, shift):
    """device_math.cuh:ecef_to_geodetic_local (angles relative to the granule centre, asin series,
    reciprocals) against the oracle's literal cartesian_to_horizontal, float64 and float32."""
    n = 20000
    g = torch.Generator().manual_seed(3)
    lat = lat0 + (torch.rand(n, generator=g, dtype=torch.float64) - 0.5) * dlat
    lon = (lon0 + (torch.rand(n, generator=g, dtype=torch.float64) - 0.5) * dlon + 180) % 360 - 180
    alt = torch.rand(n, generator=g, dtype=torch.float64) * 20000
    xyz = torch.stack(geodesy.geodetic_to_ecef(lat, lon, alt), -1)
    offset = xyz.mean(0)
    scale = float((xyz - offset).abs().max())
    pn = (xyz - offset) / scale
    lon_s = lon % 360 - 180 if shift else lon
    fr = geodesy.HorizontalFrame(scale=scale, offset=tuple(offset.tolist()), lat_min=float(lat.min()),
                                 lat_range=float(lat.max() - lat.min()), lon_min=float(lon_s.min()),
                                 lon_range=float(lon_s.max() - lon_s.min()), origin_height=20000.0, shift_lon=shift)
    hc, frs = _hc(), _frame_struct(fr)
    p64 = pn.numpy().copy()
    out64 = np.zeros_like(p64)
    hc.hc_preprocess_f64(C.byref(frs), _p(p64), _p(out64), C.c_int64(n))
    want64 = geodesy.preprocess_horizontal(pn[None], fr)[0].numpy()
    assert np.abs(out64 - want64).max() < 1e-11
    p32 = pn.float().numpy().copy()
    out32 = np.zeros_like(p32)
    hc.hc_preprocess_f32(C.byref(frs), _p(p32), _p(out32), C.c_int64(n))
    want32 = geodesy.preprocess_horizontal(pn.float()[None], fr)[0].numpy()
    assert np.abs(out32 - want32).max() < 1.5e-7 and (out32 != want32).mean() < 1e-3


@pytest.mark.parametrize("dims,key", [(3, "encoding"), (2, "surface_encoding"), (4, "encoding")])
def test_host_build_of_hash_indexing_is_bit_exact(dims, key):
    from atmonr.native import lib as L
    cfg = ngp_config()["instant_ngp"][key]
    if dims == 2:
        cfg = cfg["nested"][0]
    grid = L.grid_layout(dims, cfg)
    orc = tcnn_spec.HashGrid(dims, cfg)
    g = torch.Generator().manual_seed(4)
    x = torch.rand(3000, dims, generator=g)
    x[:6] = torch.tensor([0.0, 1.0, 0.5, 0.999999, 1e-7, 0.125])[:, None]
    xn = x.numpy().copy()
    idx = np.zeros((3000, 16, 1 << dims), np.uint32)
    w = np.zeros((3000, 16, 1 << dims), np.float32)
    _hc().hc_hashgrid_indices(C.byref(grid), _p(xn), dims, C.c_int64(3000), _p(idx), _p(w))
    assert np.array_equal(idx.astype(np.int64), orc.all_indices(x).numpy())
    for lvl in (0, 7, 15):
        _, wo = tcnn_spec.grid_corner_indices(x, orc.levels, lvl)
        assert np.array_equal(w[:, lvl], wo.numpy())


def test_philox_uniforms():
    out = np.zeros((64, 256), np.float32)
    _hc().hc_philox(C.c_uint64(123), C.c_uint64(0), C.c_int64(64), 256, _p(out))
    assert out.min() >= 0 and out.max() < 1
    assert abs(out.mean() - 0.5) < 0.01 and abs(out.var() - 1 / 12) < 0.005
    assert len(np.unique(out)) > 0.99 * out.size
    out2 = np.zeros((32, 256), np.float32)
    _hc().hc_philox(C.c_uint64(123), C.c_uint64(32), C.c_int64(32), 256, _p(out2))
    assert np.array_equal(out2, out[32:])   # the draw depends only on (seed, global ray, bin)


def test_pipeline_surface_on_cpu():
    """Construction, state-dict layout and optimizer groups of the pipelines need no GPU."""
    from helpers import FakeDataset
    from atmonr.pipelines.factory import get_pipeline
    from atmonr.utils import load_config
    cfg = load_config(os.path.join(ROOT, "configs", "instant_ngp.json"))
    scene = tiny_scene(h=3, w=3, n_views=3)
    pipe = get_pipeline(cfg["pipeline"], FakeDataset(scene))
    sd = pipe.state_dict()
    assert list(sd) == ["pos_encoder", "pos_mlp", "dir_encoder", "dir_mlp", "surf_encoder", "surf_mlp"]
    sizes = {k: v["params"].numel() for k, v in sd.items()}
    assert sizes == {"pos_encoder": 42283392, "pos_mlp": 1536, "dir_encoder": 0, "dir_mlp": 2560,
                     "surf_encoder": 5522000, "surf_mlp": 3072}
    assert pipe.fused_state is not None
    opt = pipe.get_optimizer(cfg["trainer"]["optimizer"])
    assert [g["weight_decay"] for g in opt.param_groups] == [0, 1e-2]
    assert sum(p.numel() for g in opt.param_groups for p in g["params"]) == 47812560
    pipe.load_state_dict(sd); pipe.eval(); pipe.train()
    with pytest.raises(NotImplementedError):
        get_pipeline({"type": "nope"}, FakeDataset(scene))
    ncfg = load_config(os.path.join(ROOT, "configs", "nerf.json"))
    npipe = get_pipeline(ncfg["pipeline"], FakeDataset(scene))
    assert set(npipe.state_dict()) == {"coarse", "fine"}
    assert npipe.nerf["coarse"].fc1.in_features == 76 and npipe.nerf["fine"].fc9.out_features == 260


def test_trainer_lookahead_pairs():
    """Trainer announces the NEXT batch to the pipeline one step early (Pipeline.prefetch)."""
    from atmonr.trainer import _with_lookahead
    assert list(_with_lookahead([])) == []
    assert list(_with_lookahead(["a"])) == [("a", None)]
    assert list(_with_lookahead(iter("abc"))) == [("a", "b"), ("b", "c"), ("c", None)]


# ---------------------------------------------------------------- ray table (csrc/ray_setup.cuh)
def _hc_get_rays(lat, lon, alt, thetav, phiv, height=20000.0, tol=10.0, max_iters=20):
    arrs = [np.ascontiguousarray(a, dtype=np.float32).ravel() for a in (lat, lon, alt, thetav, phiv)]
    n = arrs[0].size
    o, d, ln = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.float32)
    hc = _hc()
    hc.hc_get_rays.restype = C.c_int
    iters = hc.hc_get_rays(*map(_p, arrs), C.c_int64(n), C.c_float(height), C.c_double(tol), max_iters, _p(o), _p(d), _p(ln))
    return o, d, ln, iters


def test_host_build_of_ray_setup_matches_reference_vectors():
    """The code of atmonr_get_rays (csrc/ray_setup.cuh, built for the host) against the rays the
    REFERENCE's get_rays produced (tests/golden/reference_vectors.npz). glibc's float32/float64
    sin/cos/atan2 are the ones torch-CPU used when the vectors were made, so the agreement is to the
    last bit except where a fused multiply-add inside torch's matmul rounds differently."""
    G = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
    o, d, ln, iters = _hc_get_rays(G["rays_lat"], G["rays_lon"], G["rays_alt"], G["rays_thetav"], G["rays_phiv"])
    assert 0 <= iters <= 20
    assert np.abs(d - G["rays_dir"]).max() <= 1.2e-7          # one float32 ulp of a unit vector's component
    assert np.abs(ln - G["rays_len"]).max() <= 4e-3           # float32 ulp at 3e4 m is 2e-3
    assert np.abs(o - G["rays_origin"]).max() <= 0.5          # float32 ulp at 6.4e6 m is 0.5
    assert (o == G["rays_origin"]).mean() > 0.9 and (ln == G["rays_len"]).mean() > 0.9


def test_host_build_of_ray_setup_matches_oracle_chunk_semantics():
    """Same, against the oracle on a HARP2-shaped chunk, including the chunk-wide refinement rule:
    a steep view forces extra refinements that the nadir rays of the same chunk must take too."""
    scene_rng = np.random.default_rng(3)
    p, a = 40, 9
    lat = (30 + 5 * scene_rng.random((p, 1)) + np.zeros((1, a))).astype(np.float32)
    lon = (179.0 + 2 * scene_rng.random((p, 1)) + np.zeros((1, a))).astype(np.float32)   # across the dateline
    lon = np.where(lon > 180, lon - 360, lon).astype(np.float32)
    alt = (scene_rng.random((p, a)) * 3000).astype(np.float32)
    thetav = np.abs(np.linspace(-60, 60, a, dtype=np.float32))[None, :] + scene_rng.random((p, a)).astype(np.float32)
    phiv = (scene_rng.random((p, a)) * 360 - 180).astype(np.float32)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    for sl in (slice(None), slice(0, 1)):   # the whole chunk; one pixel alone (fewer refinements)
        args = [x[sl] for x in (lat, lon, alt, thetav, phiv)]
        want_o, want_d, want_l = geodesy.build_rays(*map(t, args), 20000.0)
        o, d, ln, iters = _hc_get_rays(*args)
        assert np.abs(d - want_d.numpy()).max() <= 1.2e-7
        assert (np.abs(ln - want_l.numpy()) / want_l.numpy()).max() <= 4e-7   # three float32 ulps (60 degree views: 4e4 m)
        assert np.abs(o - want_o.numpy()).max() <= 0.5
    # the shell is reached within the tolerance: |altitude(origin) - H| <= tol (+ float32 rounding of the origin)
    la, lo, al = geodesy.ecef_to_geodetic(*(torch.from_numpy(o.astype(np.float64))[:, k] for k in range(3)))
    assert float((al - 20000.0).abs().max()) <= 10.0 + 1.0
    # NaN geometry stays NaN (filtered later by filter_rays) and does not stall the loop
    lat2 = lat.copy(); lat2[0, 0] = np.nan
    o, d, ln, iters = _hc_get_rays(lat2, lon, alt, thetav, phiv)
    assert np.isnan(o[0]).all() and np.isfinite(o[1:]).all() and iters <= 20
    # max_iters = 0: the first guess is returned
    _, _, ln0, it0 = _hc_get_rays(lat, lon, alt, thetav, phiv, max_iters=0)
    assert it0 == 0
    want0 = ((20000.0 - alt) / np.cos(np.deg2rad(thetav.astype(np.float64)))).ravel()
    assert np.abs(ln0 - want0).max() <= 2e-2


def _hc_normalize_rays(o, d, ln):
    """wgs_84.py:316-339 the way atmonr.native.ops.normalize_rays runs it: box (native code), the two
    float64 lines of :336-337 (torch), normalisation (native code)."""
    hc = _hc()
    o, d, ln = (np.ascontiguousarray(x, dtype=np.float32) for x in (o, d, ln))
    n = ln.shape[0]
    hi_lo = np.zeros(6, np.float32)
    hc.hc_ray_extent(_p(o), _p(d), _p(ln), C.c_int64(n), _p(hi_lo))
    hi, lo = torch.from_numpy(hi_lo[:3]).double(), torch.from_numpy(hi_lo[3:]).double()
    scale = ((hi - lo).max() / 2).item()
    offset = ((hi + lo) / 2).contiguous()
    out = np.zeros_like(o)
    hc.hc_normalize_origins(_p(o), C.c_int64(n), _p(offset.numpy()), C.c_double(scale), _p(out))
    return out, scale, offset, hi_lo


def test_host_build_of_filter_and_normalize_rays_matches_reference_vectors():
    """SURVEY 8a row a2 (wgs_84.py:293-339): the per-ray functions of k_filter_rays / k_ray_extent_* /
    k_normalize_origins, compiled for the host, against the reference's own normalize_rays outputs
    (golden) and the oracle: mask, scale, offset and normalised origins BIT FOR BIT."""
    G = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
    out, scale, offset, _ = _hc_normalize_rays(G["rays_origin"], G["rays_dir"], G["rays_len"])
    assert scale == float(G["rays_scale"]) and np.array_equal(offset.numpy(), G["rays_offset"])
    assert np.array_equal(out, G["rays_origin_norm"])
    # a HARP2-shaped table with NaN pixels, against the oracle
    rng = np.random.default_rng(11)
    n = 5000
    o = (rng.standard_normal((n, 3)) * 2e5 + np.array([1.3e6, -5.0e6, 3.6e6])).astype(np.float32)
    d = rng.standard_normal((n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    ln = (2e4 + 2e4 * rng.random(n)).astype(np.float32)
    rad = rng.random(n).astype(np.float32)
    o[7, 1] = np.nan; d[90, 2] = np.nan; rad[300] = np.nan; o[4000] = np.nan; rad[4000] = np.nan
    valid = np.zeros(n, np.uint8)
    _hc().hc_filter_rays(_p(o), _p(d), _p(rad), C.c_int64(n), _p(valid))
    want_valid = geodesy.valid_ray_mask(*map(torch.from_numpy, (o, d, rad))).numpy()
    assert np.array_equal(valid.astype(bool), want_valid) and valid.sum() == n - 4
    keep = valid.astype(bool)
    out, scale, offset, hi_lo = _hc_normalize_rays(o[keep], d[keep], ln[keep])
    want, want_scale, want_offset = geodesy.normalize_rays(*map(torch.from_numpy, (o[keep], d[keep], ln[keep])))
    assert scale == want_scale and torch.equal(offset, want_offset) and np.array_equal(out, want.numpy())
    assert out.min() >= -1 and out.max() <= 1
    # an unfiltered NaN poisons its axis like torch.max / torch.min; the other axes keep their box
    _, _, _, hi_lo_nan = _hc_normalize_rays(o, d, ln)
    assert np.isnan(hi_lo_nan[[1, 4]]).all()          # y: o[7,1]; x and z hold NaN as well (o[4000], d[90,2])
    assert np.isnan(hi_lo_nan).all()
    o2 = o[keep].copy(); o2[5, 2] = np.nan
    _, _, _, hl = _hc_normalize_rays(o2, d[keep], ln[keep])
    assert np.isnan(hl[[2, 5]]).all() and np.array_equal(hl[[0, 1, 3, 4]], hi_lo[[0, 1, 3, 4]])


def test_host_build_of_normalize_rays_edge_cases():
    """Ragged and degenerate ray tables through the kernel code (host build) against the oracle, bit for bit:
    one ray, a zero-length ray alone (scale 0: the reference's 0/0 -> NaN), identical rays, huge and tiny
    coordinates, tables whose size is not a multiple of anything."""
    rng = np.random.default_rng(5)
    cases = []
    for n in (1, 2, 3, 31, 33, 257, 1000):
        o = (rng.standard_normal((n, 3)) * 10.0 ** rng.integers(-3, 7)).astype(np.float32)
        d = rng.standard_normal((n, 3)).astype(np.float32)
        ln = (rng.random(n) * 10.0 ** rng.integers(-2, 5)).astype(np.float32)
        cases.append((o, d, ln))
    one = np.array([[1.0, 2.0, 3.0]], np.float32)
    cases.append((one, np.array([[0.0, 0.0, 1.0]], np.float32), np.array([0.0], np.float32)))      # scale == 0
    cases.append((np.repeat(one, 5, 0), np.zeros((5, 3), np.float32), np.ones(5, np.float32)))      # one point, five times
    cases.append((np.array([[3e38, -3e38, 0.0], [-3e38, 3e38, 1.0]], np.float32), np.zeros((2, 3), np.float32), np.ones(2, np.float32)))
    for o, d, ln in cases:
        out, scale, offset, _ = _hc_normalize_rays(o, d, ln)
        want, want_scale, want_offset = geodesy.normalize_rays(*map(torch.from_numpy, (o, d, ln)))
        assert (scale == want_scale) or (scale != scale and want_scale != want_scale)
        assert np.array_equal(offset.numpy(), want_offset.numpy(), equal_nan=True)
        assert np.array_equal(out, want.numpy(), equal_nan=True), (o.shape, scale)
    # the zero-scale table really is the NaN case of the reference
    out, scale, _, _ = _hc_normalize_rays(*cases[7])
    assert scale == 0.0 and np.isnan(out).all()


# ---------------------------------------------------------------- dense layer on tcgen05 (csrc/linear_tc.cu)
def _bf16_round(a):
    """float32 -> nearest-even bfloat16, returned as float32 (numpy has no bfloat16)."""
    b = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    b = ((b + 0x7FFF + ((b >> 16) & 1)) >> 16) << 16
    return b.astype(np.uint32).view(np.float32)


def test_bf16_triple_split_is_float32_accurate():
    """The arithmetic of atmonr_linear_fwd_tc, emulated: three bfloat16 terms per float32 operand,
    the six partial products hh, hm, mh, mm, hl, lh accumulated in float32 (here float64, to isolate
    the error of the split itself). Single-pass bf16 is ~3e-3, the six-product form float32-grade."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((64, 332)).astype(np.float32)
    w = (rng.standard_normal((48, 332)) / 18).astype(np.float32)

    def split(v):
        hi = _bf16_round(v)
        mid = _bf16_round(v - hi)
        lo = _bf16_round(v - hi - mid)
        assert np.all((v - hi) - mid == v - hi - mid)          # the residuals are exact in float32
        return hi.astype(np.float64), mid.astype(np.float64), lo.astype(np.float64)

    (xh, xm, xl), (wh, wm, wl) = split(x), split(w)
    assert np.abs(x - (xh + xm + xl)).max() <= 2.0 ** -24 * np.abs(x).max()
    want = x.astype(np.float64) @ w.astype(np.float64).T
    six = xh @ wh.T + xh @ wm.T + xm @ wh.T + xm @ wm.T + xh @ wl.T + xl @ wh.T
    one = xh @ wh.T
    scale = np.abs(want).max()
    assert np.abs(six - want).max() / scale <= 2e-7
    assert np.abs(one - want).max() / scale >= 1e-4
    f32 = (x @ w.T).astype(np.float64)                          # a float32 GEMM, for scale
    assert np.abs(six - want).max() <= 4 * np.abs(f32 - want).max() + 1e-7 * scale


def test_linear_tc_host_side():
    from atmonr.native import lib as L, ops
    assert ops.LINEAR_TERMS in (2, 3)
    assert ops.linear_planes_bytes(256, 256, 3) == 8 * 3 * 16384
    assert ops.linear_planes_bytes(257, 76, 3) == 2 * 3 * 3 * 16384
    assert ops.linear_planes_bytes(4, 128, 2) == 4 * 2 * 16384
    with pytest.raises(L.NativeLibraryError):
        ops.linear_forward(torch.zeros(4, 8), torch.zeros(3, 8), None, False)
    with pytest.raises(L.NativeLibraryError, match="act must be"):
        L.call("atmonr_linear_fwd_tc", None, 8, None, 0, 0, None, 0, None, None, 4, 3, 8, 2, 3, None, 0, None, None, None, 3, None)
    with pytest.raises(L.NativeLibraryError, match="terms must be"):
        L.call("atmonr_linear_fwd_tc", None, 8, None, 0, 0, None, 0, None, None, 4, 3, 8, 0, 4, None, 0, None, None, None, 3, None)
    with pytest.raises(L.NativeLibraryError, match="null pointer"):
        L.call("atmonr_linear_fwd_tc", None, 8, None, 0, 0, None, 0, None, None, 4, 3, 8, 0, 2, None, 0, None, None, None, 3, None)
    with pytest.raises(L.NativeLibraryError, match="null pointer"):
        L.call("atmonr_linear_dw_tc", None, 3, None, 0, None, 8, None, 0, 0, 4, 3, 8, 3, None, None, None)
    L.call("atmonr_linear_dw_tc", None, 3, None, 0, None, 8, None, 0, 0, 0, 3, 8, 2, None, None, None)             # no rows
    with pytest.raises(L.NativeLibraryError, match="null pointer"):
        L.call("atmonr_linear_prep", None, 3, 8, 0, 3, None, None)
    with pytest.raises(L.NativeLibraryError, match="terms must be"):
        L.call("atmonr_linear_dw_tc", None, 3, None, 0, None, 8, None, 0, 0, 4, 3, 8, 1, None, None, None)
    L.call("atmonr_linear_fwd_tc", None, 8, None, 0, 0, None, 0, None, None, 0, 3, 8, 0, 3, None, 0, None, None, None, 3, None)   # no rows: nothing to do


def test_nerf_model_default_path_is_unchanged_on_cpu(monkeypatch):
    """The tensor-core layers only take CUDA activations; the CPU model (used by nothing in the product,
    but by this suite's shape checks) keeps its torch layers under either setting of DENSE_IMPL."""
    from atmonr.models.nerf import get_model
    torch.manual_seed(0)
    coarse, fine = get_model(32, 4, [4, 4, 2], 2, False)
    x = torch.rand(10, 2 * 10 + 12)
    coarse.eval()
    want = coarse(x)
    import atmonr.models.nerf as mn
    assert mn.DENSE_IMPL == "tc"
    monkeypatch.setattr(mn, "DENSE_IMPL", "library")
    got = coarse(x)
    assert torch.equal(want[0], got[0]) and torch.equal(want[1], got[1])


@pytest.mark.parametrize("with_frame", [True, False])
def test_host_build_of_nerf_point_encoder_matches_oracle_forward_and_gradient(with_frame):
    """csrc/nerf_points.cuh (atmonr_nerf_encode / _bwd, host build): rows [PE(preprocess(o + d z)) | PE(d)]
    against the oracle's preprocessing + encoders, and dL/dz (positional-encoding derivative, analytic
    float64 geodetic Jacobian, projection on the direction) against autograd through the oracle's float64
    expressions of wgs_84.py:56-97 (tolerance 2e-4 of the gradient scale: float32 sums of terms up to
    2^13 pi large, the Jacobian itself agrees to 2e-7)."""
    from oracle import nerf as onerf
    from atmonr.native import lib as L
    scene = tiny_scene()
    b = take(scene.batch, slice(0, 40))
    n, lx, ld = 12, [10, 10, 6], 4
    g = torch.Generator().manual_seed(5)
    z = (torch.rand(40, n, generator=g) * b["len"][:, None]).contiguous()
    z[0, 0] = b["len"][0] * 1.5        # below the ellipsoid: altitude clipped, its gradient masked
    fr = _frame_struct(scene.frame) if with_frame else L.disabled_frame()
    width = 2 * sum(lx) + 6 * ld
    o, d = b["origin"].numpy().copy(), b["dir"].numpy().copy()
    x = np.zeros((40 * n, width), np.float32)
    pn = np.zeros((40 * n, 3), np.float32)
    freqs = (C.c_int32 * 3)(*lx)
    hc = _hc()
    hc.hc_nerf_encode(C.byref(fr), _p(o), _p(d), _p(z.numpy()), C.c_int64(40), n, freqs, ld, _p(x), width, _p(pn))

    zt = z.clone().requires_grad_()
    pts = b["origin"][:, None] + b["dir"][:, None] * zt[..., None]
    ptn = geodesy.preprocess_horizontal(pts, scene.frame) if with_frame else pts
    want = torch.cat([onerf.pe_per_axis(ptn, lx).view(40 * n, -1),
                      onerf.pe_interleaved(b["dir"][:, None].repeat(1, n, 1), ld).view(40 * n, -1)], dim=1)
    assert np.abs(pn - ptn.detach().view(-1, 3).numpy()).max() < 1.5e-7
    # the highest frequencies multiply the last-bit differences of the preprocessed point by 2^9 pi
    assert np.abs(x - want.detach().numpy()).max() < 5e-4
    assert np.abs(x[:, :6] - want.detach().numpy()[:, :6]).max() < 2e-6

    gx = torch.randn(40 * n, width, generator=g)
    (want * gx).sum().backward()
    gz = np.zeros((40, n), np.float32)
    hc.hc_nerf_encode_bwd(C.byref(fr), _p(o), _p(d), _p(z.numpy()), _p(pn), _p(gx.numpy()), width, C.c_int64(40), n, freqs, _p(gz))
    ref = zt.grad.numpy()
    assert np.abs(ref).max() > 0
    assert np.abs(gz - ref).max() <= 2e-4 * np.abs(ref).max()
    if with_frame:
        assert scene.frame is not None and abs(float(ptn.detach().view(-1, 3)[0, 2])) == 1.0   # the clipped sample is in the set
