"""GPU parity of the native pipelines in configurations off the shipped configs. NeRF: a scalar
`L_x` without a point preprocessor, and `include_height`. oracle/nerf.py is pinned to the reference's
NeRFPipeline in both (tests/test_reference_interchange.py); first green on a B200 in round 2.
Instant-NGP: `extract` with include_height / multi_band_extinction (the operator-by-operator extract path,
which no earlier GPU test exercised, with this round's include_height fix)."""

import os

import pytest
import torch

from helpers import FakeDataset, take, tiny_scene, to_cuda
from oracle import nerf as onerf

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module", autouse=True)
def built():
    import __graft_entry__ as ge
    ge.build()
    assert torch.cuda.is_available()


@pytest.mark.parametrize("variant", ["int_L", "include_height"])
def test_nerf_pipeline_variants_match_oracle(monkeypatch, variant):
    from atmonr.pipelines.nerf import NeRFPipeline
    scene = tiny_scene()
    cfg = {"type": "NeRF", "include_height": variant == "include_height", "point_preprocessor": "", "num_bands": 4,
           "ray_origin_height": 20000, "sampler": {"N_c": 8, "N_f": 16}, "mlp_hidden_dim": 32,
           "encoder": {"L_x": 5, "L_d": 3} if variant == "int_L" else {"L_x": [4, 4, 4, 3], "L_d": 3}}
    orc = onerf.NeRFOracle(cfg, None, geo=(scene.scale, scene.offset, 20000.0))
    params = orc.init_params(seed=3)
    ds = FakeDataset(scene)
    ds.offset = scene.offset.cuda()
    pipe = NeRFPipeline(cfg, ds)
    pipe.load_state_dict({m: {k: v.detach().clone() for k, v in params[m].items()} for m in ("coarse", "fine")})
    pipe.send_tensors_to(0)
    pipe.eval()
    b = take(scene.batch, slice(0, 40))
    g = torch.Generator().manual_seed(8)
    u_c, u_f = torch.rand(40, 8, generator=g), torch.rand(40, 16, generator=g)
    draws = [u_c, u_f]
    real_rand = torch.rand
    monkeypatch.setattr(torch, "rand", lambda *a, **k: draws.pop(0).to(k.get("device", "cpu")) if draws else real_rand(*a, **k))
    res_o = orc.forward(b, params, u_c, u_f)
    loss_o = orc.loss(b, res_o)
    loss_o.backward()
    bc = to_cuda(b)
    res = pipe.forward(bc)
    loss = pipe.compute_loss(bc, res)
    loss.backward()
    rel = lambda a, c: float((a.detach().double().cpu() - c.detach().double()).abs().max() / (c.detach().double().abs().max() + 1e-30))
    for k in ("color_map_coarse", "color_map_fine", "weights_fine", "sigma_fine"):
        assert rel(res[k], res_o[k]) < 2e-3, k
    assert rel(loss, loss_o) < 1e-3
    for mode in ("coarse", "fine"):
        got = pipe.nerf[mode].fc1.weight.grad
        assert rel(got, params[mode]["fc1.weight"].grad) < 2e-2, mode
    with torch.no_grad():
        pts = (b["origin"].double() + b["dir"].double() * (0.5 * b["len"].double()[:, None])).contiguous()
        assert rel(pipe.extract(pts.cuda()), orc.extract(pts, params)) < 2e-3


@pytest.mark.parametrize("height,multi_band", [(True, False), (False, True), (True, True)])
def test_ngp_extract_with_optional_inputs_vs_oracle(height, multi_band):
    """`extract` in the off-default modes (instant_ngp.py:208-247 with include_height /
    multi_band_extinction): the 4-D grid needs the height column in extract as well (the oracle and the
    reference agree on this: tests/test_reference_interchange.py)."""
    from helpers import FakeDataset, load_params, ngp_config, random_params, tiny_scene
    from oracle.ngp import NGPOracle
    from atmonr.pipelines.instant_ngp import InstantNGPPipeline
    scene = tiny_scene()
    cfg = ngp_config(24)
    cfg["include_height"], cfg["multi_band_extinction"] = height, multi_band
    if height:
        cfg["point_preprocessor"] = ""
    orc = NGPOracle(cfg, None if height else scene.frame, scene.max_i, fp16=True, geo=(scene.scale, scene.offset, 20000.0))
    params = random_params(orc, seed=3, table_scale=2e3)
    ds = FakeDataset(scene)
    ds.offset = scene.offset.cuda()
    pipe = InstantNGPPipeline(cfg, ds)
    pipe.send_tensors_to(0)
    load_params(pipe, params)
    pipe.eval()
    # query points inside the atmosphere shell (on the scene's rays), like a voxel grid's
    b = scene.batch
    t = torch.rand(b["origin"].shape[0], 1, dtype=torch.float64, generator=torch.Generator().manual_seed(9))
    pts = (b["origin"].double() + b["dir"].double() * (t * b["len"].double()[:, None]))[:500].contiguous()
    assert pts.shape == (500, 3)
    want = orc.extract(pts, params).detach()
    got = pipe.extract(pts.cuda()).detach().cpu()
    assert got.shape == want.shape == (500, 4 if multi_band else 1)
    assert float((got - want).abs().max()) <= 2e-3 * float(want.abs().max() + 1e-12)
