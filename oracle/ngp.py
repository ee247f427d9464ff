"""Oracle: the Instant-NGP pipeline of the reference, restated on CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Glue restated from
src/atmonr/pipelines/instant_ngp.py:33-263 and PINNED to it: the reference's own
InstantNGPPipeline, run around a stand-in tinycudann that evaluates oracle/tcnn_spec.py, gives the
same results, loss, gradients and extract (tests/test_reference_interchange.py). The tiny-cuda-nn
modules themselves come from oracle/tcnn_spec.py (PARITY UNPINNED there); sampler / preprocessor /
renderer / losses are the pinned restatements in oracle/{sampling,geodesy,rendering}.py.

Differences from the reference that are deliberate and stated:
  * compositing and loss run in float32 (the reference inherits fp16 from tcnn's outputs,
    graphics_utils.py:28 -- SURVEY.md section 5 declares that the looser side);
  * the stratified uniforms `u` are an argument (fixed random draws for parity);
  * parameters are explicit flat float32 tensors, loadable into both sides.
"""

from __future__ import annotations

import torch

from oracle import rendering, sampling, tcnn_spec
from oracle.geodesy import HorizontalFrame, preprocess_horizontal

MODULES = ("pos_encoder", "pos_mlp", "dir_encoder", "dir_mlp", "surf_encoder", "surf_mlp")


class NGPOracle:
    def __init__(self, cfg: dict, frame: HorizontalFrame | None, max_i: float, fp16: bool = False, geo=None):
        """cfg is the "pipeline" section of configs/instant_ngp.json. `frame` is the granule frame of
        the 'horizontal' point preprocessor (None when the config has no preprocessor); `geo` =
        (scale, offset (3,) float64 tensor, ray_origin_height) is what pipeline.py:30-58 captures
        from the dataset, needed without a frame (`include_height`, and z in km)."""
        self.cfg, self.frame, self.max_i, self.fp16 = cfg, frame, float(max_i), fp16
        self.geo = geo
        assert not (cfg["include_height"] and frame is not None)  # pipeline.py:30-32
        ngp = cfg["instant_ngp"]
        self.n_sigma = cfg["num_bands"] if cfg["multi_band_extinction"] else 1
        self.n_pos = 4 if cfg["include_height"] else 3
        self.pos_encoder = tcnn_spec.make_encoding(self.n_pos, ngp["encoding"])
        self.pos_mlp = tcnn_spec.Network(self.pos_encoder.n_output_dims, 16, ngp["network"])
        self.dir_encoder = tcnn_spec.make_encoding(3 + 16 - self.n_sigma, ngp["dir_encoding"])
        self.dir_mlp = tcnn_spec.Network(self.dir_encoder.n_output_dims, cfg["num_bands"], ngp["rgb_network"])
        self.surf_encoder = tcnn_spec.make_encoding(5, ngp["surface_encoding"])
        self.surf_mlp = tcnn_spec.Network(self.surf_encoder.n_output_dims, cfg["num_bands"], ngp["surface_network"])
        self.loss_fn = rendering.LOSSES[cfg["loss"].lower()]

    def init_params(self, seed: int = 0) -> dict:
        gen = torch.Generator().manual_seed(seed)
        return {name: getattr(self, name).init_params(gen).requires_grad_() for name in MODULES}

    # instant_ngp.py:129-206
    def forward(self, batch: dict, params: dict, u: torch.Tensor | None):
        cfg = self.cfg
        n = cfg["num_samples_per_ray"]
        b = batch["origin"].shape[0]
        pts, z = sampling.sample_uniform(batch["origin"], batch["dir"], batch["len"], n, u)
        pts_surf = batch["origin"] + batch["dir"] * batch["len"][:, None]
        if self.frame is not None:
            pts = preprocess_horizontal(pts, self.frame)
        pts = (pts + 1) / 2
        pts_surf = (pts_surf + 1) / 2
        if cfg["include_height"]:  # instant_ngp.py:152-154 (applied to the [0,1] points, like the reference)
            scale, offset, height = self.geo
            pts = sampling.append_heights(pts, height, scale, offset)
        dirs = batch["dir"][:, None].repeat(1, n, 1)
        pts = torch.cat([pts[..., :2], pts[..., 2:3] / cfg["alt_compress_factor"], pts[..., 3:]], dim=-1)
        x = pts.view(b * n, self.n_pos)
        pos_out = self.pos_mlp.forward(
            self.pos_encoder.forward(x, params["pos_encoder"], self.fp16), params["pos_mlp"], self.fp16
        )
        dir_in = torch.cat([dirs.view(b * n, 3), pos_out[:, self.n_sigma :]], dim=1)
        color = self.dir_mlp.forward(
            self.dir_encoder.forward(dir_in, params["dir_encoder"], self.fp16), params["dir_mlp"], self.fp16
        ).view(b, n, cfg["num_bands"])
        surf_in = torch.cat([pts_surf[:, :2], dirs[:, 0]], dim=1)
        color_surf = self.surf_mlp.forward(
            self.surf_encoder.forward(surf_in, params["surf_encoder"], self.fp16), params["surf_mlp"], self.fp16
        )
        sigma = pos_out[:, : self.n_sigma].view(b, n, -1)
        color, color_surf, sigma = torch.relu(color), torch.relu(color_surf), torch.relu(sigma)
        c, _, w, c_atmo, c_surf = rendering.composite_with_surface(
            z * ((self.frame.scale if self.frame else self.geo[0]) / 1000), color, sigma, color_surf
        )
        return {
            "color_fine": color[:, :-1], "color_surf": color_surf, "color_map_surf": c_surf,
            "color_map_atmo": c_atmo, "sigma_fine": sigma[:, :-1], "color_map_fine": c,
            "weights_fine": w, "z_vals_fine": z, "pts01": x,
        }

    # instant_ngp.py:249-263
    def loss(self, batch, results):
        pred = rendering.band_select(results["color_map_fine"], batch["irgb_idx"])
        return self.loss_fn(pred, batch["rad"].to(pred.dtype), self.max_i)

    # instant_ngp.py:208-247
    def extract(self, pts, params):
        if self.frame is not None:
            pts = preprocess_horizontal(pts[None], self.frame)[0]
        pts = (pts + 1) / 2
        if self.cfg["include_height"]:  # instant_ngp.py:227-230
            scale, offset, height = self.geo
            pts = sampling.append_heights(pts[None], height, scale, offset)[0]
        pts = torch.cat([pts[..., :2], pts[..., 2:3] / self.cfg["alt_compress_factor"], pts[..., 3:]], dim=-1)
        out = self.pos_mlp.forward(
            self.pos_encoder.forward(pts.float(), params["pos_encoder"], self.fp16), params["pos_mlp"], self.fp16
        )
        return torch.clip(out[:, : self.n_sigma], min=0)

    # instant_ngp.py:107-127 + trainer.py:103-105
    def make_optimizer(self, params, opt_cfg):
        enc = [params[k] for k in ("pos_encoder", "dir_encoder", "surf_encoder") if params[k].numel()]
        mlp = [params[k] for k in ("pos_mlp", "dir_mlp", "surf_mlp")]
        kw = dict(lr=opt_cfg["lr"], betas=tuple(opt_cfg["betas"]), eps=opt_cfg["eps"])
        return torch.optim.AdamW(
            [{"params": enc, "weight_decay": 0}, {"params": mlp, "weight_decay": opt_cfg["weight_decay"]}], **kw
        )

    def train_step(self, batch, params, opt, u):
        res = self.forward(batch, params, u)
        loss = self.loss(batch, res)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss.detach(), res
