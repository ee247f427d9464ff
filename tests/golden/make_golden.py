"""Generate golden vectors by running the REFERENCE's own Python code on CPU.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py
Outputs small .npz fixtures next to this file; they are committed.  The reference modules
are imported unmodified from /root/reference/src; the four I/O-only dependencies that are
not installed (earthaccess, netCDF4, h5py, torchmetrics) are stubbed in sys.modules -- none
of them is touched by the functions called here.
"""

import sys
import types
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference/src")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    sys.path.insert(0, str(REF))
    _stub("earthaccess")
    _stub("netCDF4", Dataset=object, Variable=object)
    _stub("h5py")
    _stub("torchmetrics")
    _stub("torchmetrics.functional")
    _stub(
        "torchmetrics.functional.image",
        peak_signal_noise_ratio=None,
        structural_similarity_index_measure=None,
    )


def synthetic_geometry(h=6, w=5, n_views=7, seed=0):
    """A tiny HARP2-shaped geometry: (P, A) float32 lat/lon/alt/thetav/phiv."""
    rng = np.random.default_rng(seed)
    lat = np.linspace(35.0, 34.6, h, dtype=np.float32)[:, None] + np.zeros((1, w), np.float32)
    lon = np.linspace(-75.0, -74.5, w, dtype=np.float32)[None, :] + np.zeros((h, 1), np.float32)
    ang = np.linspace(-44.0, 44.0, n_views, dtype=np.float32)
    p = h * w
    lat = np.repeat(lat.reshape(p, 1), n_views, 1)
    lon = np.repeat(lon.reshape(p, 1), n_views, 1)
    alt = (rng.random((p, 1)) * 300).astype(np.float32) + np.zeros((1, n_views), np.float32)
    thetav = np.abs(ang)[None, :] + (rng.random((p, n_views)) * 0.5).astype(np.float32)
    phiv = np.where(ang[None, :] < 0, 180.0, 0.0).astype(np.float32) + (
        rng.random((p, n_views)) * 4 - 2
    ).astype(np.float32)
    return [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)) for a in (lat, lon, alt, thetav, phiv)]


def main():
    import_reference()
    from atmonr import encoders, graphics_utils, losses, samplers
    from atmonr.datasets.harp2 import HARP2Dataset
    from atmonr.geospatial import wgs_84
    from atmonr.models.nerf import get_model
    from atmonr.pipelines.nerf import NeRFPipeline

    out = {}

    # ---- geodesy -------------------------------------------------------------------------
    g = torch.Generator().manual_seed(1)
    lat = (torch.rand(64, generator=g, dtype=torch.float64) * 170 - 85)
    lon = (torch.rand(64, generator=g, dtype=torch.float64) * 358 - 179)
    alt = (torch.rand(64, generator=g, dtype=torch.float64) * 21000 - 500)
    x, y, z = wgs_84.horizontal_to_cartesian(lat, lon, alt)
    la2, lo2, al2 = wgs_84.cartesian_to_horizontal(x, y, z)
    out["geo_lla"] = torch.stack([lat, lon, alt], 1).numpy()
    out["geo_xyz"] = torch.stack([x, y, z], 1).numpy()
    out["geo_lla_back"] = torch.stack([la2, lo2, al2], 1).numpy()

    # ---- rays ----------------------------------------------------------------------------
    lat_g, lon_g, alt_g, thv, phv = synthetic_geometry()
    o, d, ln = wgs_84.get_rays(lat_g, lon_g, alt_g, thv, phv, 20000.0)
    on, scale, offset = wgs_84.normalize_rays(o, d, ln)
    for k, v in dict(lat=lat_g, lon=lon_g, alt=alt_g, thetav=thv, phiv=phv, origin=o, dir=d, len=ln, origin_norm=on).items():
        out["rays_" + k] = v.numpy()
    out["rays_scale"] = np.float64(scale)
    out["rays_offset"] = offset.numpy()

    # ---- the 'horizontal' point preprocessor (closure built by the reference itself) ---------
    fake_ds = SimpleNamespace(lat=lat_g, lon=lon_g, scale=scale, offset=offset, config={"ray_origin_height": 20000})
    prep = HARP2Dataset.get_point_preprocessor(fake_ds, "horizontal")
    batch = {"origin": on[::3].contiguous(), "dir": d[::3].contiguous(), "len": (ln / scale)[::3].contiguous()}
    nb = 16
    torch.manual_seed(1234)
    pts, zv = samplers.sample_uniform_bins(batch, nb)
    torch.manual_seed(1234)
    u = torch.rand((batch["origin"].shape[0], nb))
    out["samp_u"], out["samp_pts"], out["samp_z"] = u.numpy(), pts.numpy(), zv.numpy()
    pts_mid, z_mid = samplers.sample_uniform_bins(batch, nb, random=False)
    out["samp_pts_mid"], out["samp_z_mid"] = pts_mid.numpy(), z_mid.numpy()
    out["prep_f32"] = prep(pts).numpy()
    p64 = pts.double().view(-1, 3)
    out["prep_f64"] = prep(p64[None])[0].numpy()

    # the dateline branch of the closure
    lon_dl = torch.where(lon_g > -74.75, lon_g - 105.1, lon_g + 254.9)  # -> around +-180
    fake_dl = SimpleNamespace(lat=lat_g, lon=lon_dl, scale=scale, offset=offset, config={"ray_origin_height": 20000})
    prep_dl = HARP2Dataset.get_point_preprocessor(fake_dl, "horizontal")
    o_dl, d_dl, l_dl = wgs_84.get_rays(lat_g, lon_dl, alt_g, thv, phv, 20000.0)
    on_dl, sc_dl, off_dl = wgs_84.normalize_rays(o_dl, d_dl, l_dl)
    fake_dl.scale, fake_dl.offset = sc_dl, off_dl
    prep_dl = HARP2Dataset.get_point_preprocessor(fake_dl, "horizontal")
    b_dl = {"origin": on_dl[::5].contiguous(), "dir": d_dl[::5].contiguous(), "len": (l_dl / sc_dl)[::5].contiguous()}
    p_dl, _ = samplers.sample_uniform_bins(b_dl, 8, random=False)
    out["dl_lon"], out["dl_scale"], out["dl_offset"] = lon_dl.numpy(), np.float64(sc_dl), off_dl.numpy()
    out["dl_pts"], out["dl_prep"] = p_dl.numpy(), prep_dl(p_dl).numpy()

    # ---- compositing + losses ------------------------------------------------------------
    g = torch.Generator().manual_seed(2)
    zr = torch.sort(torch.rand(5, 12, generator=g) * 25, dim=1)[0]
    col = torch.rand(5, 12, 4, generator=g)
    sg1 = torch.rand(5, 12, 1, generator=g) * 0.3
    sg4 = torch.rand(5, 12, 4, generator=g) * 0.3
    cs = torch.rand(5, 4, generator=g)
    out["r_z"], out["r_col"], out["r_sg1"], out["r_sg4"], out["r_cs"] = [t.numpy() for t in (zr, col, sg1, sg4, cs)]
    for tag, sg in (("1", sg1), ("4", sg4)):
        c, a, w = graphics_utils.render(zr, col, sg)
        out[f"r_c{tag}"], out[f"r_a{tag}"], out[f"r_w{tag}"] = c.numpy(), a.numpy(), w.numpy()
        c, a, w, ca, csf = graphics_utils.render_with_surface(zr, col, sg, cs)
        out[f"rs_c{tag}"], out[f"rs_ca{tag}"], out[f"rs_cs{tag}"] = c.numpy(), ca.numpy(), csf.numpy()
    pred = torch.rand(33, generator=g) * 0.4
    gt = torch.rand(33, generator=g) * 0.4
    out["l_pred"], out["l_gt"] = pred.numpy(), gt.numpy()
    for name in ("dark", "hdr", "l1", "l1_plus_hdr", "mse", "mse_plus_hdr"):
        out["l_" + name] = getattr(losses, name + "_loss")(pred, gt, 0.37).numpy()

    # ---- positional encoding ---------------------------------------------------------------
    pe_in = torch.rand(3, 4, 3, generator=g) * 2 - 1
    out["pe_in"] = pe_in.numpy()
    out["pe_list"] = encoders.positional_encoding(pe_in, [14, 14, 10]).numpy()
    out["pe_int"] = encoders.positional_encoding(pe_in, 4).numpy()

    # ---- sample_pdf ------------------------------------------------------------------------
    wts = torch.rand(7, 16, 1, generator=g)
    zc = torch.sort(torch.rand(7, 16, generator=g), dim=1)[0] * batch["len"][:7, None]
    b7 = {k: v[:7] for k, v in batch.items()}
    torch.manual_seed(77)
    pts_f, z_f = samplers.sample_pdf(b7, wts, zc, n_samples=24)
    torch.manual_seed(77)
    u_f = torch.rand(7, 24)
    out["pdf_w"], out["pdf_zc"], out["pdf_u"] = wts.numpy(), zc.numpy(), u_f.numpy()
    out["pdf_pts"], out["pdf_z"] = pts_f.numpy(), z_f.numpy()

    # ---- NeRF model + pipeline (CPU reference), eval-mode forward and training loss/grads ----
    torch.manual_seed(5)
    cfg = {
        "type": "NeRF", "include_height": False, "point_preprocessor": "horizontal", "num_bands": 4,
        "ray_origin_height": 20000, "sampler": {"N_c": 8, "N_f": 16},
        "encoder": {"L_x": [14, 14, 10], "L_d": 4}, "mlp_hidden_dim": 32,
    }
    ds = SimpleNamespace(
        config={"ray_origin_height": 20000}, scale=scale, offset=offset, max_i=0.37,
        get_point_preprocessor=lambda name: prep,
    )
    pipe = NeRFPipeline(cfg, ds)
    nb_ = 6
    rb = {k: v[:nb_] for k, v in batch.items()}
    rb["rad"] = torch.rand(nb_, generator=g) * 0.3
    rb["irgb_idx"] = torch.randint(0, 4, (nb_,), generator=g)
    pipe.eval()  # no sigma noise -> deterministic given the two torch.rand draws
    torch.manual_seed(99)
    res = pipe.forward(rb)
    loss = pipe.compute_loss(rb, res)
    loss.backward()
    torch.manual_seed(99)
    u_c = torch.rand(nb_, 8)
    u_f2 = torch.rand(nb_, 16)
    out["nerf_rad"], out["nerf_irgb"] = rb["rad"].numpy(), rb["irgb_idx"].numpy()
    out["nerf_u_c"], out["nerf_u_f"] = u_c.numpy(), u_f2.numpy()
    for mode in ("coarse", "fine"):
        sd = pipe.nerf[mode].state_dict()
        for k, v in sd.items():
            out[f"nerf_{mode}_{k}"] = v.detach().numpy()
        for k in ("color_map", "weights", "z_vals", "sigma", "color"):
            out[f"nerf_{mode}_{k}_out"] = res[f"{k}_{mode}"].detach().numpy()
        out[f"nerf_{mode}_grad_fc1"] = pipe.nerf[mode].fc1.weight.grad.numpy()
        out[f"nerf_{mode}_grad_fc11"] = pipe.nerf[mode].fc11.weight.grad.numpy()
    out["nerf_loss"] = loss.detach().numpy()

    np.savez_compressed(HERE / "reference_vectors.npz", **out)
    total = sum(v.nbytes for v in out.values())
    print(f"wrote {len(out)} arrays, {total / 1024:.1f} KiB raw -> {HERE / 'reference_vectors.npz'}")


if __name__ == "__main__":
    main()
