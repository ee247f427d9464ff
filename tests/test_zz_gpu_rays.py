"""GPU parity of the dataset-side kernels (csrc/rays.cu): atmonr_get_rays against the oracle's
build_rays (pinned bit for bit to the reference's get_rays) and the host build of the same code,
atmonr_filter_rays / atmonr_ray_extent / atmonr_normalize_origins against the reference's expressions
(bit-exact), atmonr_gather_batch against torch indexing (bit-exact). The file sorts last on purpose: these two
entry points were opt-in (ATMONR_NATIVE_RAYS / ATMONR_NATIVE_GATHER) until they had been green on
a B200 once."""

import ctypes as C

import numpy as np
import pytest
import torch

from oracle import geodesy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def built():
    import __graft_entry__ as ge
    ge.build()
    assert torch.cuda.is_available()


def _geometry(p=300, a=9, seed=5, dateline=False):
    rng = np.random.default_rng(seed)
    lat = (30 + 5 * rng.random((p, 1)) + np.zeros((1, a))).astype(np.float32)
    lon = ((179.0 if dateline else -75.0) + 3 * rng.random((p, 1)) + np.zeros((1, a))).astype(np.float32)
    lon = np.where(lon > 180, lon - 360, lon).astype(np.float32)
    alt = (rng.random((p, a)) * 3000).astype(np.float32)
    thetav = (np.abs(np.linspace(-60, 60, a, dtype=np.float32))[None, :] + rng.random((p, a))).astype(np.float32)
    phiv = (rng.random((p, a)) * 360 - 180).astype(np.float32)
    return [torch.from_numpy(x) for x in (lat, lon, alt, thetav, phiv)]


@pytest.mark.parametrize("dateline", [False, True])
def test_get_rays_matches_oracle(dateline):
    from atmonr.native import ops
    args = _geometry(dateline=dateline)
    want_o, want_d, want_l = geodesy.build_rays(*args, 20000.0)
    o, d, ln = ops.get_rays(*(t.cuda() for t in args), 20000.0)
    assert 0 <= ops.get_rays.last_iters <= 20
    o, d, ln = o.cpu(), d.cpu(), ln.cpu()
    # The local-frame rotation is a float32 sinf/cosf expression (wgs_84.py:189-220 on float32 lat/lon):
    # the device's sinf/cosf are 1-2 ulp routines, glibc's are correctly rounded, so a direction
    # component may differ by a few float32 ulps (<= 5e-7); through the fixed point that moves the
    # length by len * 5e-7 * tan(theta) (3 cm of 40 km at 60 degrees) and the origin by less than its
    # own float32 ulp (0.5 m at 6.4e6 m).
    assert float((d - want_d).abs().max()) <= 5e-7
    assert float(((ln - want_l).abs() / want_l).max()) <= 3e-6
    assert float((o - want_o).abs().max()) <= 1.0
    # the upper end of every ray lies on the shell
    al = geodesy.ecef_to_geodetic(*(o.double()[:, k] for k in range(3)))[2]
    assert float((al - 20000.0).abs().max()) <= 11.0


def test_get_rays_edge_cases():
    from atmonr.native import ops
    args = _geometry(p=33, a=3)
    # empty chunk
    e = [t[:0].cuda() for t in args]
    o, d, ln = ops.get_rays(*e, 20000.0)
    assert o.shape == (0, 3) and d.shape == (0, 3) and ln.shape == (0,)
    # NaN geometry stays NaN and does not keep the chunk iterating
    args[0][0, 0] = float("nan")
    o, d, ln = ops.get_rays(*(t.cuda() for t in args), 20000.0)
    assert bool(o[0].isnan().all()) and bool(o[1:].isfinite().all()) and ops.get_rays.last_iters <= 20
    # max_iters = 0 returns the first guess; a one-ray chunk needs no more refinements than the full chunk
    args = _geometry(p=33, a=3)
    _, _, l0 = ops.get_rays(*(t.cuda() for t in args), 20000.0, max_iters=0)
    assert ops.get_rays.last_iters == 0
    want0 = (20000.0 - args[2].double()) / torch.cos(torch.deg2rad(args[3].double()))
    assert float((l0.cpu().double() - want0.flatten()).abs().max()) <= 5e-2   # float32 angle + cosf: 3e-7 of 4e4 m
    ops.get_rays(*(t.cuda() for t in args), 20000.0)
    full = ops.get_rays.last_iters
    ops.get_rays(*(t[:1, :1].cuda() for t in args), 20000.0)
    assert ops.get_rays.last_iters <= full


def test_get_rays_dispatch_builds_the_same_dataset(monkeypatch):
    """HARP2Dataset built with the ray-setup kernels (default) against the torch expressions on the GPU
    (ATMONR_NATIVE_RAYS=0)."""
    import json, os
    from helpers import ROOT
    from atmonr.datasets.factory import get_dataset
    cfg = json.load(open(os.path.join(ROOT, "configs", "instant_ngp.json")))["dataset"]
    monkeypatch.setenv("ATMONR_NATIVE_RAYS", "0")
    base = get_dataset(cfg, "synthetic:H=24,W=20,seed=2")
    monkeypatch.setenv("ATMONR_NATIVE_RAYS", "1")
    nat = get_dataset(cfg, "synthetic:H=24,W=20,seed=2")
    assert torch.equal(base.ray_filter, nat.ray_filter)
    assert float((base.ray_dir - nat.ray_dir).abs().max()) <= 5e-7
    assert float(((base.ray_len - nat.ray_len).abs() / base.ray_len).max()) <= 3e-6
    # one float32 ulp of an ECEF coordinate (0.5 m) moves the bounding box by 2e-6 of its size
    assert abs(base.scale - nat.scale) <= 1e-5 * base.scale
    assert float((base.ray_origin_norm - nat.ray_origin_norm).abs().max()) <= 3e-5


def test_filter_and_normalize_rays_are_bit_exact():
    """SURVEY 8a row a2 (wgs_84.py:293-339): atmonr_filter_rays / atmonr_ray_extent /
    atmonr_normalize_origins against the reference's torch expressions evaluated on the same device
    tensors and against the oracle on the CPU: mask, scale (a Python float), offset (float64[3]) and the
    normalised origins are EQUAL, for a table larger than one pass of the reduction grid."""
    import atmonr.geospatial.wgs_84 as W
    from atmonr.native import ops
    for n, seed in ((1, 0), (257, 1), (700_001, 2)):
        g = torch.Generator().manual_seed(seed)
        o = torch.randn(n, 3, generator=g) * 2e5 + torch.tensor([1.3e6, -5.0e6, 3.6e6])
        d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
        ln = 2e4 + 2e4 * torch.rand(n, generator=g)
        rad = torch.rand(n, generator=g)
        if n > 100:
            o[7, 1] = float("nan"); d[90, 2] = float("nan"); rad[30] = float("nan"); o[n - 1] = float("nan")
        oc, dc, lc, rc = (t.cuda() for t in (o, d, ln, rad))
        valid = ops.filter_rays(oc, dc, rc)
        assert valid.dtype == torch.bool and torch.equal(valid.cpu(), geodesy.valid_ray_mask(o, d, rad))
        assert int(valid.sum()) == (n - 4 if n > 100 else n)
        ok, dk, lk = oc[valid], dc[valid].contiguous(), lc[valid]
        got, scale, offset = ops.normalize_rays(ok, dk, lk)
        # the reference's expressions on the device (torch reductions are exact for max / min)
        ends = torch.cat([ok, ok + dk * lk[:, None]], dim=0)
        hi, lo = ends.max(dim=0)[0].double(), ends.min(dim=0)[0].double()
        want_scale, want_offset = ((hi - lo).max() / 2).item(), (hi + lo) / 2
        want = torch.clamp((ok - want_offset) / want_scale, -1, 1).float()
        assert isinstance(scale, float) and scale == want_scale
        assert offset.dtype == torch.float64 and offset.is_cuda and torch.equal(offset, want_offset)
        # (torch's CUDA division by a Python scalar multiplies by the float64 reciprocal, its CPU division
        # divides: the two float64 quotients can differ in the last bit, which survives the rounding to
        # float32 for about one value in 2^28. The kernel divides, like the CPU run the oracle is pinned to.)
        assert got.dtype == torch.float32 and float((got - want).abs().max()) <= 1.2e-7
        assert float((got != want).float().mean()) <= 1e-6
        w_cpu, s_cpu, off_cpu = geodesy.normalize_rays(ok.cpu(), dk.cpu(), lk.cpu())
        assert s_cpu == scale and torch.equal(off_cpu, offset.cpu()) and torch.equal(w_cpu, got.cpu())
        # the dispatching functions of the package take the kernels for CUDA tensors
        got2, scale2, offset2 = W.normalize_rays(ok, dk, lk)
        assert torch.equal(got2, got) and scale2 == scale and torch.equal(offset2, offset)
        assert torch.equal(W.filter_rays(oc, dc, rc), valid)
    # an unfiltered NaN poisons its axis of the box (torch.max / torch.min), scale and every origin with it
    got, scale, offset = ops.normalize_rays(oc, dc, lc)
    assert scale != scale and bool(offset.isnan().all()) and bool(got.isnan().all())
    # empty tables: the mask of nothing is empty, the box of nothing is an error (torch.max raises as well)
    e3, e1 = torch.empty(0, 3, device="cuda"), torch.empty(0, device="cuda")
    assert ops.filter_rays(e3, e3, e1).shape == (0,)
    from atmonr.native.lib import NativeLibraryError
    with pytest.raises(NativeLibraryError):
        ops.normalize_rays(e3, e3, e1)


def test_gather_batch_is_bit_exact(monkeypatch):
    import json, os
    from helpers import ROOT
    from atmonr.batch_loader import BatchLoader
    from atmonr.datasets.factory import get_dataset
    from atmonr.native import lib as L, ops
    cfg = json.load(open(os.path.join(ROOT, "configs", "instant_ngp.json")))["dataset"]
    ds = get_dataset(cfg, "synthetic:H=24,W=20,seed=2")
    r = len(ds)
    g = torch.Generator().manual_seed(0)
    for b in (0, 1, 257, 4096):
        idx = torch.randint(-r, r, (b,), generator=g).cuda()
        want = ds[idx]
        got = ops.gather_batch(ds._ray_tables(), idx)
        assert set(got) == set(want)
        for k in want:
            assert got[k].dtype == want[k].dtype and got[k].shape == want[k].shape, k
            assert torch.equal(got[k], want[k]), k
    # out-of-range index: reported, nothing read out of bounds
    ops.gather_batch.check = True
    try:
        with pytest.raises(IndexError):
            ops.gather_batch(ds._ray_tables(), torch.tensor([0, r], device="cuda"))
    finally:
        ops.gather_batch.check = False
    with pytest.raises(L.NativeLibraryError):
        ops.gather_batch({**ds._ray_tables(), "rad": ds.ray_rad.double()}, torch.tensor([0], device="cuda"))
    # the loader path: same batches with and without the fused gather
    monkeypatch.setenv("ATMONR_NATIVE_GATHER", "0")
    a = [b for b in BatchLoader(ds, batch_size=1000, shuffle=True, seed=3)]
    monkeypatch.setenv("ATMONR_NATIVE_GATHER", "1")
    c = [b for b in BatchLoader(ds, batch_size=1000, shuffle=True, seed=3)]
    assert len(a) == len(c)
    for x, y in zip(a, c):
        for k in x:
            assert torch.equal(x[k], y[k]), k


def test_positional_encoding_of_float64_points():
    """encoders.py:4-28 on float64 points (the extract path): phases in float64, one rounding to float32.
    At L = 14 a float32 phase is off by up to 2^13 * pi * 6e-8 = 1.5e-3 rad; the float64 flavour is not."""
    from oracle import nerf as onerf
    from atmonr.encoders import positional_encoding
    g = torch.Generator().manual_seed(0)
    p = torch.rand(513, 3, dtype=torch.float64, generator=g) * 2 - 1
    for L in ([14, 14, 10], 4):
        want = (onerf.pe_per_axis(p, L) if isinstance(L, list) else onerf.pe_interleaved(p, L)).float()
        got = positional_encoding(p.cuda(), L).cpu()
        assert got.dtype == torch.float32 and got.shape == want.shape
        assert float((got - want).abs().max()) <= 2e-7
        low = positional_encoding(p.float().cuda(), L).cpu()          # the float32 flavour is unchanged
        assert float((low - want).abs().max()) <= (3e-3 if isinstance(L, list) else 6e-6)
