"""GPU parity: every kernel of libatmonr_b200 (called through its C ABI via atmonr.native) against
the CPU oracle on identical inputs and identical random draws."""

import numpy as np
import pytest
import torch

from helpers import FakeDataset, load_params, ngp_config, random_params, take, tiny_scene, to_cuda
from oracle import geodesy, nerf as onerf, rendering, sampling, tcnn_spec
from oracle.ngp import NGPOracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene():
    return tiny_scene()


def _native():
    from atmonr.native import lib as L, ops
    L.load()
    return L, ops


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


# ------------------------------------------------------------------ sampler + preprocessor
def test_sample_uniform_bit_exact(scene):
    L, ops = _native()
    b = take(scene.batch, slice(0, 200))
    g = torch.Generator().manual_seed(3)
    for n in (16, 64, 100):
        u = torch.rand(200, n, generator=g)
        pts_o, z_o = sampling.sample_uniform(b["origin"], b["dir"], b["len"], n, u)
        pts, z = ops.sample_uniform(b["origin"].cuda(), b["dir"].cuda(), b["len"].cuda(), n, u=u.cuda())
        assert torch.equal(z.cpu(), z_o) and torch.equal(pts.cpu(), pts_o)
        pts_o, z_o = sampling.sample_uniform(b["origin"], b["dir"], b["len"], n, None)
        pts, z = ops.sample_uniform(b["origin"].cuda(), b["dir"].cuda(), b["len"].cuda(), n, random=False)
        assert torch.allclose(z.cpu(), z_o, rtol=0, atol=1e-7) and torch.allclose(pts.cpu(), pts_o, rtol=0, atol=1e-6)


def test_philox_sampler_statistics(scene):
    L, ops = _native()
    b = take(scene.batch, slice(0, 256))
    n = 64
    _, z = ops.sample_uniform(b["origin"].cuda(), b["dir"].cuda(), b["len"].cuda(), n, random=True, seed=7)
    _, z2 = ops.sample_uniform(b["origin"].cuda(), b["dir"].cuda(), b["len"].cuda(), n, random=True, seed=7)
    _, z3 = ops.sample_uniform(b["origin"].cuda(), b["dir"].cuda(), b["len"].cuda(), n, random=True, seed=8)
    assert torch.equal(z, z2) and not torch.equal(z, z3)
    t = (z.cpu() / b["len"][:, None]) * n - torch.arange(n)[None]   # position inside the bin
    assert (t >= -1e-4).all() and (t <= 1 + 1e-4).all()
    assert abs(float(t.mean()) - 0.5) < 0.01 and abs(float(t.var()) - 1 / 12) < 0.005
    # partition invariance: rays 128.. drawn alone with ray_index_base=128 give the same z
    _, zp = ops.sample_uniform(b["origin"][128:].cuda(), b["dir"][128:].cuda(), b["len"][128:].cuda(), n,
                               random=True, seed=7, ray_index_base=128)
    assert torch.equal(zp, z[128:])


def test_preprocess_matches_oracle(scene):
    L, ops = _native()
    b = take(scene.batch, slice(0, 300))
    pts, _ = sampling.sample_uniform(b["origin"], b["dir"], b["len"], 32, torch.rand(300, 32, generator=torch.Generator().manual_seed(1)))
    want32 = geodesy.preprocess_horizontal(pts, scene.frame)
    pre = FakeDataset(scene).get_point_preprocessor("horizontal")
    got32 = pre(pts.cuda()).cpu()
    assert got32.dtype == torch.float32
    # float32 output of a float64 computation: identical up to the rare last-bit rounding flip
    assert torch.allclose(got32, want32, rtol=0, atol=2.5e-7)
    assert (got32 != want32).float().mean() < 1e-3
    p64 = pts.double().view(-1, 3)
    want64 = geodesy.preprocess_horizontal(p64[None], scene.frame)[0]
    got64 = pre(p64.cuda()).cpu()
    assert got64.dtype == torch.float64 and torch.allclose(got64, want64, rtol=0, atol=1e-11)


def test_ngp_sample_points_matches_oracle(scene):
    L, ops = _native()
    cfg = ngp_config(64)
    orc = NGPOracle(cfg, scene.frame, scene.max_i)
    b = take(scene.batch, slice(0, 128))
    u = torch.rand(128, 64, generator=torch.Generator().manual_seed(2))
    res = orc.forward(b, orc.init_params(0), u)
    pre = FakeDataset(scene).get_point_preprocessor("horizontal")
    x01, z = ops.ngp_sample_points(pre.frame, b["origin"].cuda(), b["dir"].cuda(), b["len"].cuda(), 64, 8.0, u=u.cuda())
    assert torch.equal(z.cpu(), res["z_vals_fine"])
    assert torch.allclose(x01.cpu(), res["pts01"], rtol=0, atol=1.5e-7)


# ------------------------------------------------------------------ hash grid
@pytest.mark.parametrize("dims,key", [(3, "encoding"), (2, "surface_encoding"), (4, "encoding")])
def test_hashgrid_indices_bit_exact_and_features(dims, key):
    L, ops = _native()
    cfg = ngp_config()["instant_ngp"][key]
    if dims == 2:
        cfg = cfg["nested"][0]
    grid_o = tcnn_spec.HashGrid(dims, cfg)
    grid = L.grid_layout(dims, cfg)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(4096, dims, generator=g)
    x[:8] = torch.tensor([0.0, 1.0, 0.5, 0.25, 0.999999, 1e-7, 0.125, 0.75])[:, None]  # edges
    if dims >= 3:
        x[:, 2] *= 0.125   # compressed altitude range
    idx = ops.hashgrid_indices(grid, x.cuda()).cpu().to(torch.int64) & 0xFFFFFFFF
    assert torch.equal(idx, grid_o.all_indices(x))
    params = (torch.rand(grid_o.n_params, generator=g) * 2 - 1) * 0.5
    want = grid_o.forward(x, params, fp16=True)
    p = params.cuda().requires_grad_()
    got = ops.HashGridFn.apply(x.cuda(), p, p.detach().half(), grid)
    # interpolation = one fp16 FMA per corner in corner order (tcnn's arithmetic): reproduced exactly
    assert torch.equal(got.cpu(), want)
    want32 = grid_o.forward(x, params, fp16=False)
    assert rel_err(got, want32) < 2e-3
    # backward: gradient of sum(out * r) w.r.t. the table
    r = torch.randn(4096, grid_o.n_output_dims, generator=g)
    pc = params.clone().requires_grad_()
    (grid_o.forward(x, pc, fp16=True) * r).sum().backward()
    (got * r.cuda()).sum().backward()
    assert torch.allclose(p.grad.cpu(), pc.grad, rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------ MLPs
@pytest.mark.parametrize("n_in,n_out,key", [(32, 16, "network"), (19, 4, "rgb_network"), (36, 4, "surface_network"), (16, 4, "rgb_network")])
def test_mlp_forward_backward(n_in, n_out, key):
    L, ops = _native()
    cfg = ngp_config()["instant_ngp"][key]
    net_o = tcnn_spec.Network(n_in, n_out, cfg)
    g = torch.Generator().manual_seed(11)
    params = net_o.init_params(g)
    m = 1000  # not a multiple of the 128-row tile
    x = torch.randn(m, n_in, generator=g)
    r = torch.randn(m, n_out, generator=g)
    pc, xc = params.clone().requires_grad_(), x.clone().requires_grad_()
    want = net_o.forward(xc, pc, fp16=True)
    (want * r).sum().backward()
    shape = L.mlp_shape(n_in, n_out, cfg)
    p, xg = params.cuda().requires_grad_(), x.cuda().requires_grad_()
    got = ops.MlpFn.apply(xg, p, p.detach().half(), shape)
    (got * r.cuda()).sum().backward()
    assert rel_err(got, want) < 1e-4
    assert rel_err(p.grad, pc.grad) < 1e-4
    assert rel_err(xg.grad, xc.grad) < 1e-4
    assert rel_err(got, net_o.forward(x, params, fp16=False)) < 5e-3   # vs pure fp32


# ------------------------------------------------------------------ compositing + loss
@pytest.mark.parametrize("k,v,surf", [(4, 1, True), (4, 4, True), (4, 1, False), (2, 2, False), (3, 1, True)])
def test_composite_forward_backward(k, v, surf):
    L, ops = _native()
    g = torch.Generator().manual_seed(21)
    b, n = 37, 77   # ragged: not a multiple of the warp size
    z = torch.sort(torch.rand(b, n, generator=g) * 25, dim=1)[0]
    col = (torch.rand(b, n, k, generator=g) - 0.2)   # raw values, some negative -> exercised by relu
    sig = (torch.rand(b, n, v, generator=g) - 0.3) * 0.5
    cs = torch.rand(b, k, generator=g) - 0.1 if surf else None
    ra, rs = torch.randn(b, k, generator=g), torch.randn(b, k, generator=g)
    for relu in (True, False):
        zc, cc, sc = z.clone().requires_grad_(), col.clone().requires_grad_(), (sig if relu else sig.abs()).clone().requires_grad_()
        csc = cs.clone().requires_grad_() if surf else None
        act = torch.relu if relu else (lambda t: t)
        if surf:
            cm, al, w, ca, csf = rendering.composite_with_surface(zc, act(cc), act(sc), act(csc))
            loss = (ca * ra).sum() + (csf * rs).sum() + cm.sum()
        else:
            cm, al, w = rendering.composite(zc, act(cc), act(sc))
            loss = (cm * ra).sum()
        loss.backward()
        zg, cg, sg = z.cuda().requires_grad_(), col.cuda().requires_grad_(), (sig if relu else sig.abs()).cuda().requires_grad_()
        csg = cs.cuda().requires_grad_() if surf else None
        gm, ga, gs, gw, gal = ops.CompositeFn.apply(zg, cg, sg, csg, 1.0, relu)
        if surf:
            lg = (ga * ra.cuda()).sum() + (gs * rs.cuda()).sum() + gm.sum()
        else:
            lg = (gm * ra.cuda()).sum()
        lg.backward()
        assert rel_err(gm, cm) < 1e-5 and rel_err(gw, w) < 1e-5 and rel_err(gal, al) < 1e-5
        assert rel_err(cg.grad, cc.grad) < 1e-4
        assert rel_err(sg.grad, sc.grad) < 1e-4
        assert rel_err(zg.grad, zc.grad) < 1e-3
        if surf:
            assert rel_err(csg.grad, csc.grad) < 1e-5


def test_composite_edge_cases():
    L, ops = _native()
    # single sample per ray, zero density, saturated density
    z = torch.tensor([[0.7], [1.3]]).cuda()
    col = torch.ones(2, 1, 4).cuda()
    sig = torch.tensor([[[0.0]], [[1e4]]]).cuda()
    cm, ca, cs, w, a = ops.CompositeFn.apply(z, col, sig, torch.ones(2, 4).cuda(), 1.0, False)
    want = rendering.composite_with_surface(z.cpu(), col.cpu(), sig.cpu(), torch.ones(2, 4))
    assert torch.allclose(cm.cpu(), want[0]) and torch.allclose(w.cpu(), want[2])
    # empty batch
    e = ops.CompositeFn.apply(torch.zeros(0, 8).cuda(), torch.zeros(0, 8, 4).cuda(), torch.zeros(0, 8, 1).cuda(), None, 1.0, True)
    assert e[0].shape == (0, 4)


@pytest.mark.parametrize("kind", list(rendering.LOSSES))
def test_band_loss(kind):
    L, ops = _native()
    g = torch.Generator().manual_seed(31)
    b = 1234
    cm = (torch.rand(b, 4, generator=g) * 0.4).requires_grad_()
    band = torch.randint(0, 4, (b,), generator=g)
    rad = torch.rand(b, generator=g) * 0.4
    want = rendering.LOSSES[kind](rendering.band_select(cm, band), rad, 0.37)
    want.backward()
    cg = cm.detach().cuda().requires_grad_()
    got = ops.band_loss(cg, band.cuda(), rad.cuda(), 0.37, kind)
    got.backward()
    assert rel_err(got, want) < 1e-5
    assert rel_err(cg.grad, cm.grad) < 1e-4


# ------------------------------------------------------------------ optimizer
def test_fused_adamw_matches_torch():
    from atmonr.optim import FusedAdamW
    g = torch.Generator().manual_seed(41)
    n = 100003
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    mine = torch.nn.Parameter(p0.clone().cuda())
    kw = dict(lr=1e-2, betas=(0.9, 0.99), eps=1e-15, weight_decay=1e-2)
    o_ref = torch.optim.AdamW([ref], foreach=False, **kw)
    o_mine = FusedAdamW([mine], **kw)
    for step in range(5):
        gr = torch.randn(n, generator=g) * (0.1 if step != 2 else 0.0)   # a zero-gradient step too
        ref.grad, mine.grad = gr.clone(), gr.clone().cuda()
        o_ref.step(); o_mine.step()
        assert torch.allclose(mine.detach().cpu(), ref.detach(), rtol=1e-6, atol=1e-7), step
    from atmonr.native.modules import shadow_of
    assert torch.equal(shadow_of(mine).cpu(), mine.detach().cpu().half())
    sd = o_mine.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}


# ------------------------------------------------------------------ the whole NGP step
def _pipeline(scene, cfg):
    from atmonr.pipelines.instant_ngp import InstantNGPPipeline
    pipe = InstantNGPPipeline(cfg, FakeDataset(scene))
    pipe.send_tensors_to(0)
    return pipe


@pytest.mark.parametrize("impl", ["tc", "tc-compact-bwd", "simt", "tc-nocache", "tc-narrow-bwd"])
def test_ngp_pipeline_forward_loss_gradients_vs_oracle(scene, impl, monkeypatch):
    """impl = tc: dense layers on tcgen05 (fp16 gradient operands under a power-of-two scale), 128-row
    forward tiles, 256-row backward tiles reading the forward's cached encoding;
    tc-nocache: no encoding cache -> 128-row backward that re-gathers the table;
    tc-narrow-bwd: the 128-row backward with the cache; simt: thread-per-sample FMA kernels
    (fp32 gradients)."""
    from atmonr.native import fused
    monkeypatch.setattr(fused, "FIELD_IMPL", "simt" if impl == "simt" else "tc")
    if impl == "tc-compact-bwd":   # the backward visits only the samples that carry a gradient
        monkeypatch.setattr(fused, "COMPACT_BWD", True)
    if impl == "tc-nocache":
        monkeypatch.setattr(fused, "ENC_CACHE_BYTES", 0)
    if impl == "tc-narrow-bwd":
        monkeypatch.setenv("ATMONR_BWD_NARROW", "1")
    cfg = ngp_config(64)
    orc16 = NGPOracle(cfg, scene.frame, scene.max_i, fp16=True)
    orc32 = NGPOracle(cfg, scene.frame, scene.max_i, fp16=False)
    params = random_params(orc16, seed=0, table_scale=2e3)
    b = take(scene.batch, slice(0, 96))
    u = torch.rand(96, 64, generator=torch.Generator().manual_seed(9))
    res = orc16.forward(b, params, u)
    loss = orc16.loss(b, res)
    loss.backward()
    pipe = _pipeline(scene, cfg)
    assert pipe.fused_state is not None
    load_params(pipe, params)
    bc = to_cuda(b)
    out = pipe.forward(bc, u=u.cuda())
    lg = pipe.compute_loss(bc, out)
    lg.backward()
    for key in ("color_map_fine", "color_map_atmo", "color_map_surf"):
        assert rel_err(out[key], res[key]) < 1e-3, key                      # native rounding points emulated
    res32 = orc32.forward(b, params, u)
    assert rel_err(out["color_map_fine"], res32["color_map_fine"]) < 1e-2   # vs pure fp32 (fp16 table/MLP path)
    assert rel_err(lg, loss) < 1e-3
    for key in ("weights_fine", "sigma_fine", "color_fine", "z_vals_fine", "color_surf"):
        assert rel_err(out[key], res[key]) < 2e-3, key
    # tc: every gradient operand of the MLP backward is rounded to fp16 (relative 2^-11) once
    grad_tol = 2e-3 if impl == "simt" else 4e-3
    for name in ("pos_mlp", "dir_mlp", "surf_mlp", "pos_encoder", "surf_encoder"):
        got = getattr(pipe, name).params.grad
        assert rel_err(got, params[name].grad) < grad_tol, name
    # modular (operator-by-operator) path agrees with the fused path
    pipe.fused_state = None
    out_m = pipe.forward(bc, u=u.cuda())
    # (same encodings bit for bit; the dense layers accumulate in a different order, which can flip the
    # fp16 rounding of a hidden activation)
    assert rel_err(out_m["color_map_fine"], out["color_map_fine"]) < 1e-4


@pytest.mark.parametrize("height,multi_band", [(True, False), (False, True), (True, True)])
def test_ngp_optional_inputs_vs_oracle(scene, height, multi_band):
    """`include_height` (4-D hash grid over [x, y, z/8, h], samplers.py:168-195) and
    `multi_band_extinction` (one density per band): off in the shipped configs. Both run through the fused
    launch chain (height-column sampler, field kernels templated on the grid dimensionality and the number
    of densities, compositing with one density per band) AND through the operator-by-operator path; each
    is compared with the oracle. include_height excludes the 'horizontal' preprocessor (pipeline.py:30-32)."""
    from atmonr.native import fused
    cfg = ngp_config(24)
    cfg["include_height"], cfg["multi_band_extinction"] = height, multi_band
    if height:
        cfg["point_preprocessor"] = ""
    geo = (scene.scale, scene.offset, 20000.0)
    orc = NGPOracle(cfg, None if height else scene.frame, scene.max_i, fp16=True, geo=geo)
    params = random_params(orc, seed=3, table_scale=2e3)
    b = take(scene.batch, slice(0, 40))
    u = torch.rand(40, 24, generator=torch.Generator().manual_seed(4))
    res = orc.forward(b, params, u)
    loss = orc.loss(b, res)
    loss.backward()
    ds = FakeDataset(scene)
    ds.offset = scene.offset.cuda()
    from atmonr.pipelines.instant_ngp import InstantNGPPipeline
    pipe = InstantNGPPipeline(cfg, ds)
    pipe.send_tensors_to(0)
    st = pipe.fused_state
    assert st is not None and fused.field_impl(st) == "simt"
    assert st.n_density == (4 if multi_band else 1) and (st.height is not None) == height
    load_params(pipe, params)
    bc = to_cuda(b)
    for path in ("fused", "modular"):
        if path == "modular":
            pipe.fused_state = None
        for name in ("pos_mlp", "dir_mlp", "surf_mlp", "pos_encoder", "surf_encoder"):
            getattr(pipe, name).params.grad = None
        out = pipe.forward(bc, u=u.cuda())
        lg = pipe.compute_loss(bc, out)
        lg.backward()
        assert out["sigma_fine"].shape[-1] == (4 if multi_band else 1)
        for key in ("color_map_fine", "color_map_atmo", "color_map_surf", "sigma_fine", "color_fine"):
            assert rel_err(out[key], res[key]) < 2e-3, (path, key)
        if height:
            assert rel_err(out["norm_heights_fine"], res["pts01"][:, 3].view(40, 24)) < 1e-6, path
        assert rel_err(lg, loss) < 1e-3, path
        for name in ("pos_mlp", "dir_mlp", "surf_mlp", "pos_encoder", "surf_encoder"):
            assert rel_err(getattr(pipe, name).params.grad, params[name].grad) < 4e-3, (path, name)


def test_ngp_training_tracks_oracle(scene):
    """Same initial parameters, same batches, same uniform draws: the native loss curve follows
    the oracle's (torch AdamW on the fp16-emulating restatement)."""
    cfg = ngp_config(32)
    orc = NGPOracle(cfg, scene.frame, scene.max_i, fp16=True)
    params = random_params(orc, seed=1, table_scale=1.0)
    opt_cfg = {"lr": 1e-2, "betas": [0.9, 0.99], "eps": 1e-15, "weight_decay": 1e-2}
    opt = orc.make_optimizer(params, opt_cfg)
    pipe = _pipeline(scene, cfg)
    load_params(pipe, params)
    opt_n = pipe.get_optimizer(opt_cfg)
    g = torch.Generator().manual_seed(77)
    lo, ln = [], []
    n_rays = scene.batch["origin"].shape[0]
    for step in range(30):
        sel = torch.randperm(n_rays, generator=g)[:64]
        b = take(scene.batch, sel)
        u = torch.rand(64, 32, generator=g)
        l_o, _ = orc.train_step(b, params, opt, u)
        bc = to_cuda(b)
        out = pipe.forward(bc, u=u.cuda())
        l_n = pipe.compute_loss(bc, out)
        opt_n.zero_grad(); l_n.backward(); opt_n.step()
        lo.append(float(l_o)); ln.append(float(l_n))
    lo, ln = np.array(lo), np.array(ln)
    assert ln[-1] < ln[0]                                  # it trains
    assert np.max(np.abs(ln - lo) / lo) < 0.05, (lo, ln)   # and tracks the oracle step by step


@pytest.mark.parametrize("impl", ["tc", "simt"])
def test_extract_matches_oracle(scene, impl, monkeypatch):
    """instant_ngp.py:208-247. tc: k_extract_sigma_tc (dense layers on tcgen05, the default);
    simt: the thread-per-voxel kernel kept as its cross-check. 5000 voxels = 39 full tiles + a ragged one."""
    from atmonr.native import fused
    monkeypatch.setattr(fused, "FIELD_IMPL", impl)
    cfg = ngp_config(16)
    orc = NGPOracle(cfg, scene.frame, scene.max_i, fp16=True)
    params = random_params(orc, seed=2, table_scale=5e3)
    g = torch.Generator().manual_seed(3)
    pts = (torch.rand(5000, 3, generator=g, dtype=torch.float64) * 2 - 1) * 0.9
    want = orc.extract(pts, params)
    pipe = _pipeline(scene, cfg)
    load_params(pipe, params)
    pipe.eval()
    with torch.no_grad():
        got = pipe.extract(pts.cuda())
    assert got.shape == (5000, 1) and (got >= 0).all()
    assert rel_err(got, want) < 2e-3
    assert float((want > 0).float().mean()) > 0.05          # not trivially all zeros


# ------------------------------------------------------------------ NeRF helpers
def test_positional_encoding_and_sample_pdf():
    L, ops = _native()
    from atmonr.encoders import positional_encoding
    g = torch.Generator().manual_seed(51)
    p = torch.rand(50, 7, 3, generator=g) * 2 - 1
    got = positional_encoding(p.cuda(), [14, 14, 10]).cpu()
    want = onerf.pe_per_axis(p, [14, 14, 10])
    assert got.shape == want.shape and torch.allclose(got, want, atol=2e-3)   # sin/cos of arguments up to 2^13*pi
    assert torch.allclose(got[..., :4], want[..., :4], atol=1e-6)
    got = positional_encoding(p.cuda(), 4).cpu()
    assert torch.allclose(got, onerf.pe_interleaved(p, 4), atol=1e-5)
    w = torch.rand(40, 64, 1, generator=g)
    zc = torch.sort(torch.rand(40, 64, generator=g), dim=1)[0]
    u = torch.rand(40, 128, generator=g)
    z_o, inds_o = sampling.inverse_cdf_z(w, zc, u)
    z, inds = ops.sample_pdf_z(w[..., 0].cuda(), zc.cuda(), u.cuda())
    assert (torch.diff(z, dim=1) >= 0).all()
    assert torch.allclose(z.cpu(), z_o, atol=1e-5)
    assert (inds.cpu() == inds_o).float().mean() > 0.999      # identical except u within an ulp of a CDF edge


# ------------------------------------------------------------------ tensor-core operand layouts
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_tcgen05_probe(mode):
    L, ops = _native()
    g = torch.Generator().manual_seed(61 + mode)
    a = torch.randn(128, 32, generator=g).half()
    b = torch.randn(128, 32, generator=g).half()
    d = torch.full((128, 64 if mode >= 3 else 32), float("nan"), device="cuda")
    ag, bg = a.cuda(), b.cuda()   # keep both alive: a temporary's block would be reused
    L.call("atmonr_tc_probe", L.ptr(ag), L.ptr(bg), mode, L.ptr(d), L.stream())
    torch.cuda.synchronize()
    af, bf = a.float(), b.float()
    if mode == 0:
        want, got = af @ bf[:32].T, d.cpu()
    elif mode == 1:
        want, got = af @ bf[:32], d.cpu()
    elif mode == 2:
        want, got = af.T @ bf, d.cpu()[:32]
    elif mode == 3:   # two sample groups per MMA: the gradient is the sum of the diagonal blocks
        dc = d.cpu()
        want, got = af.T @ bf, dc[:32, :32] + dc[32:64, 32:64]
    else:
        dc = d.cpu()
        want, got = af.T @ bf[:, :16], dc[:32, :16] + dc[32:64, 16:32]
    assert torch.allclose(got, want, rtol=1e-3, atol=1e-3), float((got - want).abs().max())


@pytest.mark.parametrize("impl,n_steps", [("simt", 300), ("tc", 300), ("tc", 2000)])
def test_ngp_loss_curve_tracks_oracle(scene, impl, n_steps, monkeypatch):
    """north_star: 'loss curves tracking the reference over 2k steps'. 300 / 2000 optimisation
    steps with identical initial parameters, batches and uniform draws; small hash tables
    (log2 T = 12) keep the CPU oracle's dense AdamW cheap.

    AdamW with eps = 1e-15 (configs/instant_ngp.json:94) normalises every non-zero gradient to a
    full-size step, so trajectories of two correct implementations separate chaotically after a
    few dozen steps (the 30-step test above checks step-by-step agreement). What is checked here
    is the curve: same smoothed shape and same final loss level, tighter for the fp32-gradient SIMT
    backward than for the tcgen05 backward, whose fp16 gradient operands flush the smallest
    gradients to zero (as tiny-cuda-nn's fp16 backward does)."""
    import copy
    from atmonr.native import fused
    monkeypatch.setattr(fused, "FIELD_IMPL", impl)
    cfg = copy.deepcopy(ngp_config(32))
    cfg["instant_ngp"]["encoding"]["log2_hashmap_size"] = 12
    cfg["instant_ngp"]["surface_encoding"]["nested"][0]["log2_hashmap_size"] = 12
    orc = NGPOracle(cfg, scene.frame, scene.max_i, fp16=True)
    params = random_params(orc, seed=4, table_scale=1.0)
    opt_cfg = {"lr": 1e-2, "betas": [0.9, 0.99], "eps": 1e-15, "weight_decay": 1e-2}
    opt = orc.make_optimizer(params, opt_cfg)
    init = {k: v.detach().clone() for k, v in params.items()}
    g = torch.Generator().manual_seed(123)
    n_rays = scene.batch["origin"].shape[0]
    draws = [(torch.randperm(n_rays, generator=g)[:48], torch.rand(48, 32, generator=g)) for _ in range(n_steps)]
    lo = []
    for sel, u in draws:
        l_o, _ = orc.train_step(take(scene.batch, sel), params, opt, u)
        lo.append(float(l_o))
    lo = np.array(lo)
    win = max(40, n_steps // 10)   # the 48-ray batches make single losses noisy: compare running means
    smooth = lambda a: np.convolve(a, np.ones(win) / win, mode="valid")
    so = smooth(lo)

    def native_run():
        pipe = _pipeline(scene, cfg)
        assert pipe.fused_state is not None
        load_params(pipe, init)
        opt_n = pipe.get_optimizer(opt_cfg)
        ln = []
        for sel, u in draws:
            bc = to_cuda(take(scene.batch, sel))
            l_n = pipe.compute_loss(bc, pipe.forward(bc, u=u.cuda()))
            opt_n.zero_grad(); l_n.backward(); opt_n.step()
            ln.append(float(l_n.detach()))
        return np.array(ln)

    def check(ln):
        sn = smooth(ln)
        dev = np.abs(sn - so) / so
        print(f"[{impl}] first/last smoothed loss: oracle {so[0]:.4f}/{so[-1]:.4f} native {sn[0]:.4f}/{sn[-1]:.4f}; "
              f"max dev {dev.max():.3f} mean dev {dev.mean():.3f}; first 10 steps max dev {np.max(np.abs(ln[:10]-lo[:10])/lo[:10]):.4f}")
        print("  smoothed oracle", np.round(so[:: n_steps // 15], 4).tolist())
        print("  smoothed native", np.round(sn[:: n_steps // 15], 4).tolist())
        ratio = sn / so
        return {
            "first 10 steps agree step by step": np.max(np.abs(ln[:10] - lo[:10]) / lo[:10]) < 0.05,
            "both make the same real progress": sn[-1] < 0.2 * sn[0] and so[-1] < 0.2 * so[0],
            "same curve (chaotic bumps stay within a band)": 0.5 < ratio.min() and ratio.max() < 2.0,
            "same final loss level": 0.6 < sn[-1] / so[-1] < 1.5,
        }

    # The native trajectory is not reproducible run to run (fp32 atomics in the table-gradient
    # scatter feed a chaotic optimiser), so a band violation is re-tried once before it counts.
    for attempt in range(2):
        verdict = check(native_run())
        if all(verdict.values()):
            break
    assert all(verdict.values()), [k for k, ok in verdict.items() if not ok]


def test_prefetched_sampler_is_bit_identical(scene):
    """InstantNGPPipeline.prefetch: the next batch's sample points are computed on a side stream
    underneath the current step's backward. Two training steps with and without it must give the
    same bits: same losses, same gradients (same draw counter -> same stratified samples)."""
    cfg = ngp_config(64)
    orc = NGPOracle(cfg, scene.frame, scene.max_i, fp16=True)
    params = random_params(orc, seed=0, table_scale=2e3)
    b0, b1, b2 = (to_cuda(take(scene.batch, slice(96 * k, 96 * (k + 1)))) for k in range(3))

    def run(use_prefetch):
        pipe = _pipeline(scene, cfg)
        load_params(pipe, params)
        st = pipe.fused_state
        losses, grads = [], []
        seq = [b0, b1, b2]
        if use_prefetch:
            pipe.prefetch(seq[0])            # announced, never launched by a backward: sampled in line
        for k, b in enumerate(seq):
            if use_prefetch and k + 1 < len(seq):
                pipe.prefetch(seq[k + 1])
            out = pipe.forward(b)
            loss = pipe.compute_loss(b, out)
            for name in pipe.module_names:
                getattr(pipe, name).params.grad = None
            loss.backward()
            if use_prefetch and k + 1 < len(seq):
                assert st.pending and st.pending[0]["done"] is not None   # launched by the backward
            losses.append((loss.detach().clone(), st.last["x01"].clone(), st.last["z"].clone()))
            grads.append(pipe.pos_encoder.params.grad.detach().clone())
        torch.cuda.synchronize()
        assert not st.pending
        return losses, grads, st.step

    l_a, g_a, steps_a = run(False)
    l_b, g_b, steps_b = run(True)
    assert steps_a == steps_b == 3
    for (la, xa, za), (lb, xb, zb) in zip(l_a, l_b):
        assert torch.equal(xa, xb) and torch.equal(za, zb)
        assert rel_err(la, lb) < 1e-6
    for x, y in zip(g_a, g_b):
        # fp32 atomics reorder between runs; the sampled points are what must be identical
        assert rel_err(x, y) < 1e-5


def test_compact_backward_equals_dense_backward(scene, monkeypatch):
    """atmonr_composite_bwd_compact + atmonr_ngp_field_bwd_tc_compact: the list holds exactly the
    samples with a positive raw density, in ray order, with the dense kernel's gradients; the field
    backward over the list gives the dense backward's parameter gradients (fp32 atomics reorder)."""
    from atmonr.native import fused
    L, ops = _native()
    cfg = ngp_config(64)
    orc = NGPOracle(cfg, scene.frame, scene.max_i, fp16=True)
    params = random_params(orc, seed=1, table_scale=2e3)
    b = to_cuda(take(scene.batch, slice(0, 200)))
    u = torch.rand(200, 64, generator=torch.Generator().manual_seed(3)).cuda()
    grads = {}
    for mode in (True, False):
        monkeypatch.setattr(fused, "COMPACT_BWD", mode)
        pipe = _pipeline(scene, cfg)
        load_params(pipe, params)
        out = pipe.forward(b, u=u)
        pipe.compute_loss(b, out).backward()
        grads[mode] = {n: getattr(pipe, n).params.grad.clone() for n in ("pos_encoder", "pos_mlp", "dir_mlp", "surf_mlp")}
        if mode:
            st = pipe.fused_state
            sig = st.last["sigma_raw"]
            assert int(st.last["n_active"]) == int((sig > 0).sum())
            assert 0 < int(st.last["n_active"]) < sig.numel()    # the case is not degenerate
    for n in grads[True]:
        assert rel_err(grads[True][n], grads[False][n]) < 1e-5, n
    # operator level: list contents against the dense operator
    z, sig, col = st.last["z"], st.last["sigma_raw"], st.last["color_raw"]
    bb, nn = z.shape
    cs = st.last["color_surf_raw"]
    cmap, catmo, csurf, tsurf, _, _ = ops.composite_forward(z, col, sig, cs, st.z_scale, relu=True, want_weights=False, want_alpha=False)
    g = torch.randn(bb, 4, generator=torch.Generator().manual_seed(5)).cuda()
    dcol, dsig, dcs = ops.composite_backward(z, col, sig, cs, catmo, tsurf, g, g, st.z_scale, relu=True)
    idx, n_act, dcol_c, dsig_c, dcs_c = ops.composite_backward_compact(z, col, sig, cs, catmo, tsurf, g, g, st.z_scale, relu=True)
    n = int(n_act)
    idx = idx[:n].long()
    want_idx = torch.nonzero(sig.view(-1) > 0).view(-1)
    assert torch.equal(torch.sort(idx).values, want_idx)
    ray = idx // nn
    same_ray = ray[1:] == ray[:-1]
    assert bool((idx[1:][same_ray] > idx[:-1][same_ray]).all())      # ordered inside a ray ...
    assert int((~same_ray).sum()) + 1 == int(torch.unique(ray).numel())  # ... and a ray's samples are contiguous
    assert torch.equal(dcol_c[:n], dcol.view(-1, 4)[idx]) and torch.equal(dsig_c[:n], dsig.view(-1, 1)[idx])
    assert torch.equal(dcs_c, dcs)
    dead = torch.ones(bb * nn, dtype=torch.bool, device=z.device)
    dead[idx] = False
    assert float(dcol.view(-1, 4)[dead].abs().max()) == 0.0 and float(dsig.view(-1)[dead].abs().max()) == 0.0


@pytest.mark.parametrize("case", ["none-listed", "all-listed", "ragged"])
def test_compact_backward_edge_cases(scene, case):
    """The list is empty (every raw density <= 0: the field backward must do nothing and return),
    holds every sample (relu == 0 lists all), or ends in the middle of a 256-row tile."""
    L, ops = _native()
    import ctypes as C
    from atmonr.native import fused
    cfg = ngp_config(64)
    orc = NGPOracle(cfg, scene.frame, scene.max_i, fp16=True)
    params = random_params(orc, seed=1, table_scale=2e3)
    pipe = _pipeline(scene, cfg)
    load_params(pipe, params)
    b = to_cuda(take(scene.batch, slice(0, 37)))      # 37 x 64 = 2368 samples = 9.25 tiles
    u = torch.rand(37, 64, generator=torch.Generator().manual_seed(3)).cuda()
    out = pipe.forward(b, u=u)
    st = pipe.fused_state
    z, col, cs = st.last["z"], st.last["color_raw"], st.last["color_surf_raw"]
    sig = st.last["sigma_raw"].clone()
    relu = True
    if case == "none-listed":
        sig = -sig.abs() - 1e-3
    elif case == "all-listed":
        relu = False
    bb, nn = z.shape
    cmap, catmo, csurf, tsurf, _, _ = ops.composite_forward(z, col, sig, cs, st.z_scale, relu=relu, want_weights=False, want_alpha=False)
    g = torch.randn(bb, 4, generator=torch.Generator().manual_seed(5)).cuda()
    absmax = torch.zeros(1, device="cuda")
    idx, n_act, dcol_c, dsig_c, _ = ops.composite_backward_compact(z, col, sig, cs, catmo, tsurf, g, g, st.z_scale, relu=relu, grad_absmax=absmax)
    n = int(n_act)
    assert n == {"none-listed": 0, "all-listed": bb * nn}.get(case, int((sig > 0).sum()))
    if case == "ragged":
        assert n % 256 != 0
    # the field backward over the list against the dense field backward on the scattered gradients
    x01 = st.last["x01"]
    dense_col = torch.zeros(bb * nn, 4, device="cuda")
    dense_sig = torch.zeros(bb * nn, device="cuda")
    dense_col[idx[:n].long()] = dcol_c[:n]
    dense_sig[idx[:n].long()] = dsig_c[:n, 0]
    t16, pw16, dw16 = pipe.pos_encoder.table_f16(), pipe.pos_mlp.weights_f16(), pipe.dir_mlp.weights_f16()
    _, _, enc = fused.field_forward(st, t16, pw16, dw16, x01, b["dir"].contiguous(), bb, nn, want_enc=True)
    outs = {}
    for mode in ("compact", "dense"):
        d_t = torch.zeros(pipe.pos_encoder.params.numel(), device="cuda")
        d_pw = torch.zeros(pipe.pos_mlp.params.numel(), device="cuda")
        d_dw = torch.zeros(pipe.dir_mlp.params.numel(), device="cuda")
        if mode == "compact":
            L.call("atmonr_ngp_field_bwd_tc_compact", C.byref(st.grid3), C.byref(st.pos_mlp), L.ptr(pw16), C.byref(st.dir_mlp),
                   L.ptr(dw16), L.ptr(x01), L.ptr(b["dir"].contiguous()), L.ptr(enc), L.ptr(idx), L.ptr(n_act), L.ptr(dsig_c),
                   L.ptr(dcol_c), L.ptr(absmax), bb, nn, L.ptr(d_t), L.ptr(d_pw), L.ptr(d_dw), L.stream())
        else:
            L.call("atmonr_ngp_field_bwd_tc", C.byref(st.grid3), L.ptr(t16), C.byref(st.pos_mlp), L.ptr(pw16), C.byref(st.dir_mlp),
                   L.ptr(dw16), L.ptr(x01), L.ptr(b["dir"].contiguous()), L.ptr(enc), L.ptr(dense_sig), L.ptr(dense_col),
                   L.ptr(absmax), bb, nn, L.ptr(d_t), L.ptr(d_pw), L.ptr(d_dw), L.stream())
        torch.cuda.synchronize()
        outs[mode] = (d_t, d_pw, d_dw)
    for a, c in zip(outs["compact"], outs["dense"]):
        if case == "none-listed":
            assert float(a.abs().max()) == 0.0 and float(c.abs().max()) == 0.0
        else:
            assert rel_err(a, c) < 1e-5
