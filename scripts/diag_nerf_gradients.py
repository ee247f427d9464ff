"""NeRF gradients against oracle/nerf.py at configs/nerf.json size, per loss term (coarse only / fine only / both):
locates which gradient path a mismatch comes from (the fine -> coarse path is rounding-dominated)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atmospheric-neural-rendering_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from helpers import FakeDataset, take, tiny_scene, to_cuda
from oracle import nerf as onerf, rendering
from atmonr.pipelines.nerf import NeRFPipeline
from atmonr.native import ops
scene = tiny_scene()
cfg = json.load(open(os.path.join(ROOT, "configs", "nerf.json")))["pipeline"]
if len(sys.argv) > 1:
    cfg["mlp_hidden_dim"] = int(sys.argv[1])
n_c, n_f = cfg["sampler"]["N_c"], cfg["sampler"]["N_f"]
orc = onerf.NeRFOracle(cfg, scene.frame)
params = orc.init_params(seed=3)
b = take(scene.batch, slice(0, 40))
g = torch.Generator().manual_seed(8)
u_c, u_f = torch.rand(40, n_c, generator=g), torch.rand(40, n_f, generator=g)
rel = lambda a, c: float((a.detach().double().cpu() - c.detach().double()).abs().max() / (c.detach().double().abs().max() + 1e-30))
rel2 = lambda a, c: float((a.detach().double().cpu() - c.detach().double()).norm() / (c.detach().double().norm() + 1e-30))
real_rand = torch.rand
for which in ("coarse_only", "fine_only", "both"):
    for m in params.values():
        for v in m.values():
            v.grad = None
    res_o = orc.forward(b, params, u_c, u_f)
    band, rad = b["irgb_idx"], b["rad"]
    lc = ((rendering.band_select(res_o["color_map_coarse"], band) - rad) ** 2).mean()
    lf = ((rendering.band_select(res_o["color_map_fine"], band) - rad) ** 2).mean()
    lo = {"coarse_only": lc, "fine_only": lf, "both": lc + lf}[which]
    lo.backward()
    pipe = NeRFPipeline(cfg, FakeDataset(scene))
    pipe.load_state_dict({m: {k: v.detach().clone() for k, v in params[m].items()} for m in ("coarse", "fine")})
    pipe.send_tensors_to(0)
    pipe.eval()
    draws = [u_c, u_f]
    torch.rand = lambda *a, **k: draws.pop(0).to(k.get("device", "cpu")) if draws else real_rand(*a, **k)
    bc = to_cuda(b)
    res = pipe.forward(bc)
    torch.rand = real_rand
    l1 = ops.band_loss(res["color_map_coarse"], bc["irgb_idx"], bc["rad"], 1.0, "mse")
    l2 = ops.band_loss(res["color_map_fine"], bc["irgb_idx"], bc["rad"], 1.0, "mse")
    ln = {"coarse_only": l1, "fine_only": l2, "both": l1 + l2}[which]
    ln.backward()
    out = {"loss": rel(ln, lo), "cmap_c": rel(res["color_map_coarse"], res_o["color_map_coarse"]), "cmap_f": rel(res["color_map_fine"], res_o["color_map_fine"]), "z_f": rel(res["z_vals_fine"], res_o["z_vals_fine"])}
    for mode in ("coarse", "fine"):
        for name in ("fc1.weight", "fc6.weight", "fc9.weight", "fc11.weight"):
            layer, attr = name.split(".")
            got = getattr(getattr(pipe.nerf[mode], layer), attr).grad
            want = params[mode][name].grad
            if got is None or want is None:
                out[f"{mode}.{name}"] = None
            else:
                out[f"{mode}.{name}"] = (round(rel(got, want), 4), round(rel2(got, want), 4), float(want.abs().max()))
    print(which, json.dumps(out))
