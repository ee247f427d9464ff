"""NeRF pipeline (reference: src/atmonr/pipelines/nerf.py): coarse + fine hierarchical sampling
with sinusoidal positional encoding and the AtmoNeRF MLPs.

Every stage is a kernel of libatmonr_b200, forward and backward: stratified sampler, inverse-CDF
sampling + sort (atmonr_sample_pdf_train / _bwd), point on the ray + geodetic preprocessing +
positional encodings of point and direction in one pass (atmonr_nerf_encode / _bwd), the eleven
dense layers on tcgen05 (models/nerf.py), compositing (atmonr_composite_fwd / _bwd / _bwd_weights).
The reference keeps the gradient path fine loss -> fine sample distances -> coarse weights alive
(samplers.py:96 only detaches the bin width): here it runs through atmonr_nerf_encode_bwd (encoding
derivative + float64 geodetic Jacobian), atmonr_composite_dz, atmonr_sample_pdf_bwd and
atmonr_composite_bwd_weights instead of torch graphs. `include_height`, an integer L_x or a
preprocessor without a frame description take the operator-by-operator path below.
"""

from __future__ import annotations

from itertools import chain
from typing import Any, Mapping

import torch
import torch.nn.functional as F
from torch.optim import Adam, Optimizer

from atmonr.encoders import positional_encoding
from atmonr.geospatial.wgs_84 import cartesian_to_horizontal
from atmonr.graphics_utils import render
from atmonr.models.nerf import get_model
from atmonr.native import lib as L, ops
from atmonr.pipelines.pipeline import Pipeline
from atmonr.samplers import append_heights, inverse_cdf_z, sample_pdf, sample_uniform_bins


class NeRFPipeline(Pipeline):
    def __init__(self, config: dict, dataset) -> None:
        super().__init__(config, dataset)
        coarse, fine = get_model(
            hidden_dim=config["mlp_hidden_dim"], N_lambda=config["num_bands"], L_x=config["encoder"]["L_x"],
            L_d=config["encoder"]["L_d"], include_height=config["include_height"],
        )
        self.nerf = {"coarse": coarse, "fine": fine}
        self.training = True

    def send_tensors_to(self, device: int) -> None:
        self.device = device
        for mode in self.nerf:
            self.nerf[mode] = self.nerf[mode].to(device)

    def get_optimizer(self, config: dict) -> Optimizer:
        params = chain(self.nerf["coarse"].parameters(), self.nerf["fine"].parameters())
        return Adam(params=params, lr=config["lr"])

    # -------------------------------------------------------------------------------------
    def _preprocess(self, pts: torch.Tensor) -> torch.Tensor:
        if not self.point_preprocessor:
            return pts
        frame = getattr(self.point_preprocessor, "frame", None)
        if frame is None or not (torch.is_grad_enabled() and pts.requires_grad):
            return self.point_preprocessor(pts)
        # differentiable evaluation of harp2.py:372-386 (fine pass only)
        off = torch.tensor([frame.offset[0], frame.offset[1], frame.offset[2]], dtype=torch.float64, device=pts.device)
        xyz = pts * frame.scale + off
        lat, lon, alt = cartesian_to_horizontal(xyz[..., 0], xyz[..., 1], xyz[..., 2])
        if frame.shift_lon:
            lon = lon % 360 - 180
        lat = 2 * (lat - frame.lat_min) / frame.lat_range - 1
        lon = 2 * (lon - frame.lon_min) / frame.lon_range - 1
        alt = 2 * alt / frame.origin_height - 1
        return torch.clip(torch.stack([lat, lon, alt], dim=-1).to(pts.dtype), min=-1, max=1)

    def _fused_frame(self):
        """Frame description for atmonr_nerf_encode, or None when the configuration needs the modular path."""
        cfg = self.config
        l_x = cfg["encoder"]["L_x"]
        if cfg["include_height"] or not (isinstance(l_x, list) and len(l_x) == 3) or not isinstance(cfg["encoder"]["L_d"], int):
            return None
        if not self.point_preprocessor:
            return L.disabled_frame()
        return getattr(self.point_preprocessor, "frame", None)

    def _forward(self, mode: str, ray_batch, weights_coarse=None, z_vals_coarse=None):
        """nerf.py:73-167."""
        assert (mode == "coarse") == (z_vals_coarse is None)
        cfg = self.config
        b = ray_batch["origin"].shape[0]
        l_x, l_d = cfg["encoder"]["L_x"], cfg["encoder"]["L_d"]
        frame = self._fused_frame()
        pts = None
        if frame is not None:
            # sample distances -> [encoded point | encoded direction] rows, one kernel each way
            if mode == "coarse":
                n = cfg["sampler"]["N_c"]
                z_vals = sample_uniform_bins(ray_batch, n_bins=n)[1]
            else:
                n = cfg["sampler"]["N_c"] + cfg["sampler"]["N_f"]
                u = torch.rand((b, cfg["sampler"]["N_f"]), device=z_vals_coarse.device)
                z_vals = inverse_cdf_z(weights_coarse[..., 0], z_vals_coarse, u)
            x = ops.NerfEncodeFn.apply(z_vals, ray_batch["origin"], ray_batch["dir"], frame, tuple(l_x), l_d)[0]
        else:
            if mode == "coarse":
                n = cfg["sampler"]["N_c"]
                pts, z_vals = sample_uniform_bins(ray_batch, n_bins=n)
            else:
                n = cfg["sampler"]["N_c"] + cfg["sampler"]["N_f"]
                pts, z_vals = sample_pdf(ray_batch, weights_coarse, z_vals_coarse, n_samples=cfg["sampler"]["N_f"])
            pts = self._preprocess(pts)
            if cfg["include_height"]:
                pts = append_heights(pts, self.ray_origin_height, self.scale, self.offset)
            pts_enc = positional_encoding(pts, l_x).view((b * n, -1))
            dirs = ray_batch["dir"][:, None].repeat(1, n, 1)
            dirs_enc = positional_encoding(dirs, l_d).view((b * n, -1))
            x = torch.cat([pts_enc, dirs_enc], dim=1)
        color, sigma = self.nerf[mode](x)
        color = color.view(b, n, -1)
        sigma = sigma.view(b, n, 1 if mode == "coarse" else -1)
        color = torch.exp(torch.clamp(color, max=11))  # nerf.py:150 (after the sigmoid)
        sigma = F.relu(sigma)
        z_km = z_vals * (self.scale / 1000)
        color_map, _, weights = render(z_km, color, sigma)   # the weights stay differentiable (CompositeFn)
        results = {f"color_{mode}": color, f"sigma_{mode}": sigma, f"color_map_{mode}": color_map,
                   f"weights_{mode}": weights, f"z_vals_{mode}": z_vals}
        if cfg["include_height"]:
            results[f"norm_heights_{mode}"] = pts[..., 3]
        return results

    def forward(self, ray_batch: Mapping[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        results = self._forward("coarse", ray_batch)
        results.update(self._forward("fine", ray_batch, weights_coarse=results["weights_coarse"],
                                     z_vals_coarse=results["z_vals_coarse"]))
        # keys the shipped Trainer reads from every pipeline (trainer.py:129-137); the reference
        # NeRF pipeline lacks them, here they are defined sensibly
        results["color_map_atmo"] = results["color_map_fine"]
        results["color_map_surf"] = torch.zeros_like(results["color_map_fine"])
        return results

    def extract(self, pts: torch.Tensor) -> torch.Tensor:
        """nerf.py:190-217."""
        if self.point_preprocessor:
            pts = self.point_preprocessor(pts[None])[0]
        if self.config["include_height"]:
            pts = append_heights(pts[None], self.ray_origin_height, self.scale, self.offset)[0]
        # float64 points (scripts/extract.py) are encoded in float64 and rounded once, like the reference
        enc = positional_encoding(pts, self.config["encoder"]["L_x"]).view(pts.shape[0], -1).float()
        _, sigma = self.nerf["fine"].forward_pos_only(enc)
        return torch.clip(sigma, min=0)

    def compute_loss(self, ray_batch, results) -> torch.Tensor:
        """nerf.py:219-240: un-normalised MSE of the coarse and of the fine rendering."""
        band, rad = ray_batch["irgb_idx"], ray_batch["rad"]
        return (ops.band_loss(results["color_map_coarse"], band, rad, 1.0, "mse")
                + ops.band_loss(results["color_map_fine"], band, rad, 1.0, "mse"))

    def state_dict(self) -> Mapping[str, Mapping[str, Any]]:
        return {mode: self.nerf[mode].state_dict() for mode in ("coarse", "fine")}

    def load_state_dict(self, state_dict: dict) -> None:
        for mode in ("coarse", "fine"):
            self.nerf[mode].load_state_dict(state_dict[mode])

    def train(self) -> None:
        self.training = True
        for net in self.nerf.values():
            net.train()

    def eval(self) -> None:
        self.training = False
        for net in self.nerf.values():
            net.eval()
