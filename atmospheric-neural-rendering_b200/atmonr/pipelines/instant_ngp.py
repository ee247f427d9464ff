"""Instant-NGP pipeline (reference: src/atmonr/pipelines/instant_ngp.py).

Same plugin surface (constructor, module names, state-dict layout, result keys, AdamW parameter
groups). The six tiny-cuda-nn modules are replaced by atmonr.native.modules.{Encoding,Network};
the whole forward/backward runs through the fused launch chain of libatmonr_b200
(atmonr.native.fused; `include_height` and `multi_band_extinction` included), and through the modular
operators chained exactly like the reference's forward for shapes the chain does not cover.
"""

from __future__ import annotations

from itertools import chain
from typing import Any, Mapping

import torch
import torch.nn.functional as F
from torch.optim import Optimizer

from atmonr.graphics_utils import render_with_surface
from atmonr.native import fused, lib as L, ops
from atmonr.native.modules import Encoding, Network
from atmonr.optim import FusedAdamW
from atmonr.pipelines.pipeline import Pipeline
from atmonr.samplers import append_heights, sample_uniform_bins

MODULE_NAMES = ["pos_encoder", "pos_mlp", "dir_encoder", "dir_mlp", "surf_encoder", "surf_mlp"]


class InstantNGPPipeline(Pipeline):
    def __init__(self, config: dict, dataset) -> None:
        super().__init__(config, dataset)
        self.num_density_outputs = config["num_bands"] if config["multi_band_extinction"] else 1
        num_inputs = 4 if config["include_height"] else 3
        ngp = config["instant_ngp"]
        self.module_names = list(MODULE_NAMES)
        self.pos_encoder = Encoding(num_inputs, ngp["encoding"])
        self.pos_mlp = Network(self.pos_encoder.n_output_dims, 16, ngp["network"])
        self.dir_encoder = Encoding(3 + 16 - self.num_density_outputs, ngp["dir_encoding"])
        self.dir_mlp = Network(self.dir_encoder.n_output_dims, config["num_bands"], ngp["rgb_network"])
        self.surf_encoder = Encoding(2 + 3, ngp["surface_encoding"])
        self.surf_mlp = Network(self.surf_encoder.n_output_dims, config["num_bands"], ngp["surface_network"])
        self.training = True
        self.max_i = dataset.max_i
        self.loss_name = config["loss"].lower()
        if self.loss_name not in ops.LOSS_KINDS:
            raise KeyError(self.loss_name)
        self.fused_state = self._make_fused_state()
        # data-parallel bookkeeping (Trainer sets it under torchrun): the stratified draws are keyed by
        # (seed, draw counter, GLOBAL ray index in the batch, bin), so a batch gets the same draws
        # however it is sharded; shards are equal-sized (BatchLoader), hence base = rank * local rays
        self.rank, self.world_size = 0, 1

    # -------------------------------------------------------------------------------------
    def _make_fused_state(self):
        """The fused launch chain covers: 3-D grid (or 4-D with `include_height`), 16 levels x 2 features,
        one density (or four with `multi_band_extinction`), 4 bands, pos_mlp 32->[32]->16, dir_mlp
        19->[32,32]->4 (16->[32,32]->4 with four densities), surf 2-D grid + SH -> [32,32] -> 4, and either the
        'horizontal' preprocessor or none. The shipped shape runs on the tcgen05 kernels, the two optional
        inputs on the thread-per-sample kernels of the same chain (atmonr.native.fused.field_impl)."""
        cfg = self.config
        frame = getattr(self.point_preprocessor, "frame", None)
        nd = self.num_density_outputs
        dir_in = (32, 2) if nd == 1 else (16, 2)
        ok = (
            nd in (1, 4) and cfg["num_bands"] == 4
            and (self.point_preprocessor is None or frame is not None)
            and self.pos_encoder.grid is not None and self.pos_encoder.grid.n_levels == 16
            and len(self.pos_encoder.parts) == 1
            and self.surf_encoder.grid is not None and self.surf_encoder.grid.n_levels == 16
            and [p[0] for p in self.surf_encoder.parts] == ["HashGrid", "SphericalHarmonics"]
            and [p[0] for p in self.dir_encoder.parts] == ["SphericalHarmonics", "Identity"]
            and (self.pos_mlp.shape.in_pad, self.pos_mlp.shape.n_hidden) == (32, 1)
            and (self.dir_mlp.shape.in_pad, self.dir_mlp.shape.n_hidden) == dir_in
            and (self.surf_mlp.shape.in_pad, self.surf_mlp.shape.n_hidden) == (48, 2)
        )
        if not ok:
            return None
        height = None
        if cfg["include_height"]:   # samplers.py:168-195 on the [0,1]^3 point (instant_ngp.py:155-156)
            height = (float(self.scale), [float(v) for v in self.offset.tolist()], float(self.ray_origin_height))
        return fused.NGPState(
            frame=frame if frame is not None else L.disabled_frame(),
            grid3=self.pos_encoder.grid, grid2=self.surf_encoder.grid,
            pos_mlp=self.pos_mlp.shape, dir_mlp=self.dir_mlp.shape, surf_mlp=self.surf_mlp.shape,
            n_samples=int(cfg["num_samples_per_ray"]), alt_compress=float(cfg["alt_compress_factor"]),
            z_scale=self.scale / 1000, n_density=nd, height=height,
        )

    def send_tensors_to(self, device: int) -> None:
        self.device = device
        for name in self.module_names:
            getattr(self, name).to(device)

    def get_optimizer(self, config: dict) -> Optimizer:
        """instant_ngp.py:107-127: AdamW, weight decay on the MLPs only. All three encoder parameters
        are listed, the empty one of dir_encoder (SH + identity have no weights) included, exactly like
        the reference (tcnn exposes a 0-element Parameter): the groups are [0,1,2] / [3,4,5], so
        optimizer state dicts interchange with reference checkpoints. FusedAdamW.step skips it."""
        enc = chain(self.pos_encoder.parameters(), self.dir_encoder.parameters(), self.surf_encoder.parameters())
        mlp = chain(self.pos_mlp.parameters(), self.dir_mlp.parameters(), self.surf_mlp.parameters())
        return FusedAdamW(
            [{"params": list(enc), "weight_decay": 0},
             {"params": list(mlp), "weight_decay": config["weight_decay"]}],
            **config,
        )

    # -------------------------------------------------------------------------------------
    def forward(self, ray_batch: Mapping[str, torch.Tensor], u: torch.Tensor | None = None) -> dict[str, torch.Tensor]:
        """`u` (B, N) optionally fixes the stratified draws (parity tests); by default they come
        from the in-kernel counter-based generator, a fresh stream every call."""
        if self.fused_state is None:
            return self._forward_modular(ray_batch, u)
        st = self.fused_state
        st.ray_index_base = self.rank * ray_batch["origin"].shape[0]
        shadows = (self.pos_encoder.table_f16(), self.pos_mlp.weights_f16(), self.dir_mlp.weights_f16(),
                   self.surf_encoder.table_f16(), self.surf_mlp.weights_f16())
        cmap, catmo, csurf = fused.NGPRenderFn.apply(
            self.pos_encoder.params, self.pos_mlp.params, self.dir_mlp.params, self.surf_encoder.params,
            self.surf_mlp.params, st, shadows, ray_batch["origin"], ray_batch["dir"], ray_batch["len"], u)
        return fused.LazyResults(
            {"color_map_fine": cmap, "color_map_atmo": catmo, "color_map_surf": csurf}, st)

    def prefetch(self, ray_batch: Mapping[str, torch.Tensor]) -> None:
        """Optional: announce the NEXT training batch before running the current step. Its sample
        points (the parameter-independent, FP64-bound part of forward) are then computed on a side
        stream underneath the current step's backward (atmonr.native.fused.schedule_prefetch).
        forward() picks them up when it is given the same `origin` tensor; results are those of
        an in-line sampler call with the same draw counter."""
        if self.fused_state is not None and self.training:
            self.fused_state.ray_index_base = self.rank * ray_batch["origin"].shape[0]
            fused.schedule_prefetch(self.fused_state, ray_batch["origin"], ray_batch["dir"], ray_batch["len"])

    def _forward_modular(self, ray_batch, u=None):
        """instant_ngp.py:129-206 operator by operator (non-default configurations)."""
        cfg = self.config
        b, n = ray_batch["origin"].shape[0], cfg["num_samples_per_ray"]
        if u is None:
            pts, z_vals = sample_uniform_bins(ray_batch, n)
        else:
            pts, z_vals = ops.sample_uniform(ray_batch["origin"], ray_batch["dir"], ray_batch["len"], n, u=u)
        pts_surf = ray_batch["origin"] + ray_batch["dir"] * ray_batch["len"][:, None]
        if self.point_preprocessor:
            pts = self.point_preprocessor(pts)
        pts, pts_surf = (pts + 1) / 2, (pts_surf + 1) / 2
        if cfg["include_height"]:
            pts = append_heights(pts, self.ray_origin_height, self.scale, self.offset)
        dirs = ray_batch["dir"][:, None].repeat(1, n, 1)
        pts = torch.cat([pts[..., :2], pts[..., 2:3] / cfg["alt_compress_factor"], pts[..., 3:]], dim=-1)
        pos_out = self.pos_mlp(self.pos_encoder(pts.view(b * n, -1)))
        nd = self.num_density_outputs
        color = self.dir_mlp(self.dir_encoder(torch.cat([dirs.view(b * n, 3), pos_out[:, nd:]], dim=1)))
        color = color.view(b, n, cfg["num_bands"])
        color_surf = self.surf_mlp(self.surf_encoder(torch.cat([pts_surf[:, :2], dirs[:, 0]], dim=1)))
        sigma = pos_out[..., :nd].view(b, n, -1)
        color, color_surf, sigma = F.relu(color), F.relu(color_surf), F.relu(sigma)
        cmap, _, weights, catmo, csurf = render_with_surface(z_vals * (self.scale / 1000), color, sigma, color_surf)
        results = {
            "color_fine": color[:, :-1], "color_surf": color_surf, "color_map_surf": csurf,
            "color_map_atmo": catmo, "sigma_fine": sigma[:, :-1], "color_map_fine": cmap,
            "weights_fine": weights, "z_vals_fine": z_vals,
        }
        if cfg["include_height"]:
            results["norm_heights_fine"] = pts[..., 3]
        return results

    def extract(self, pts: torch.Tensor) -> torch.Tensor:
        """instant_ngp.py:208-247: (n,3) normalised scene points -> (n, n_density) extinction."""
        st = self.fused_state
        if st is not None and st.n_density == 1 and st.height is None:
            return fused.extract_sigma(st, self.pos_encoder.table_f16(), self.pos_mlp.weights_f16(), pts)
        if self.point_preprocessor:
            pts = self.point_preprocessor(pts[None])[0]
        pts = (pts + 1) / 2
        if self.config["include_height"]:  # instant_ngp.py:227-230: the 4-D grid needs its height column here too
            pts = append_heights(pts[None], self.ray_origin_height, self.scale, self.offset)[0]
        pts = torch.cat([pts[..., :2], pts[..., 2:3] / self.config["alt_compress_factor"], pts[..., 3:]], dim=-1)
        out = self.pos_mlp(self.pos_encoder(pts.float()))
        return torch.clip(out[..., : self.num_density_outputs], min=0)

    def set_draw_counter(self, n: int) -> None:
        """Resume: continue the stratified-draw stream after `n` training steps instead of replaying it."""
        if self.fused_state is not None:
            self.fused_state.step = int(n)

    def compute_loss(self, ray_batch, results) -> torch.Tensor:
        """instant_ngp.py:249-263, one kernel: band select + loss + d(loss)/d(colour map)."""
        return ops.band_loss(results["color_map_fine"], ray_batch["irgb_idx"], ray_batch["rad"], self.max_i, self.loss_name)

    # -------------------------------------------------------------------------------------
    def state_dict(self) -> Mapping[str, Mapping[str, Any]]:
        return {name: getattr(self, name).state_dict() for name in self.module_names}

    def load_state_dict(self, state_dict: dict) -> None:
        for name in self.module_names:
            getattr(self, name).load_state_dict(state_dict[name])

    def train(self) -> None:
        self.training = True
        for name in self.module_names:
            getattr(self, name).train()

    def eval(self) -> None:
        self.training = False
        for name in self.module_names:
            getattr(self, name).eval()
