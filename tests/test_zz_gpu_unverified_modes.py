"""GPU parity of the native NeRF pipeline in the configurations off the shipped config: a scalar
`L_x` without a point preprocessor, and `include_height`. oracle/nerf.py is pinned to the reference's
NeRFPipeline in both (tests/test_reference_interchange.py); the native side of these modes has not run
on a B200 yet, so the tests are gated like tests/test_zz_gpu_linear_tc.py (ATMONR_RUN_UNVERIFIED=1)."""

import os

import pytest
import torch

from helpers import FakeDataset, take, tiny_scene, to_cuda
from oracle import nerf as onerf

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("ATMONR_RUN_UNVERIFIED") != "1",
                                 reason="not validated on hardware yet (set ATMONR_RUN_UNVERIFIED=1)")]


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
