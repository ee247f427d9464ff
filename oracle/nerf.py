"""Oracle: the reference's NeRF path (positional encoding, AtmoNeRF MLP, coarse+fine pipeline).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PINNED against tests/golden (vectors made by
the reference's own NeRFPipeline on CPU).  Reference: src/atmonr/encoders.py:4-28,
src/atmonr/models/nerf.py:6-144, src/atmonr/pipelines/nerf.py:73-240.
"""

from __future__ import annotations

import math

import torch

from oracle import rendering, sampling
from oracle.geodesy import HorizontalFrame, preprocess_horizontal


def pe_per_axis(p, freqs_per_axis):
    """encoders.py:21-27 (list-L variant): per axis [sin(2^l pi p) for l | cos(...) for l]."""
    outs = []
    for axis, n in enumerate(freqs_per_axis):
        f = (2.0 ** torch.linspace(0, n - 1, steps=n)) * torch.pi
        arg = f[..., None, :] * p[..., axis, None]
        outs += [torch.sin(arg), torch.cos(arg)]
    return torch.cat(outs, dim=-1)


def pe_interleaved(p, n):
    """encoders.py:14-20 (int-L variant): returns (M, C, 2n) with [sin, cos] interleaved per
    frequency."""
    flat = p.reshape(-1, p.shape[-1])
    f = (2.0 ** torch.linspace(0, n - 1, steps=n)) * torch.pi
    arg = f[None, None, :] * flat[..., None]
    return torch.stack([torch.sin(arg), torch.cos(arg)], dim=-1).reshape(flat.shape[0], flat.shape[1], 2 * n)


LAYERS = tuple(f"fc{i}" for i in range(1, 12))


def mlp_shapes(pos_ch, dir_ch, out_ch, vol_ch, hidden):
    """models/nerf.py:33-43: (out, in) of fc1..fc11."""
    h = hidden
    return [
        (h, pos_ch), (h, h), (h, h), (h, h), (h, h), (h, h + pos_ch), (h, h), (h, h),
        (h + vol_ch, h), (h // 2, h + dir_ch), (out_ch, h // 2),
    ]


def init_mlp(shapes, gen):
    """nn.Linear default bias init + kaiming_normal_(mode='fan_out') weights (nerf.py:45-46)."""
    sd = {}
    for name, (o, i) in zip(LAYERS, shapes):
        sd[name + ".weight"] = torch.randn(o, i, generator=gen) * math.sqrt(2.0 / o)
        bound = 1 / math.sqrt(i)
        sd[name + ".bias"] = (torch.rand(o, generator=gen) * 2 - 1) * bound
    return sd


def mlp_forward(sd, x_pos, x_dir, hidden, noise=None):
    """models/nerf.py:48-93.  noise: optional (M, V) N(0,1) draws added to sigma (training)."""
    lin = lambda n, t: torch.nn.functional.linear(t, sd[n + ".weight"], sd[n + ".bias"])
    h = torch.relu(lin("fc1", x_pos))
    for n in ("fc2", "fc3", "fc4", "fc5"):
        h = torch.relu(lin(n, h))
    h = torch.relu(lin("fc6", torch.cat([h, x_pos], dim=1)))
    h = torch.relu(lin("fc7", h))
    h = torch.relu(lin("fc8", h))
    h = lin("fc9", h)
    sigma = h[:, hidden:]
    if noise is not None:
        sigma = sigma + noise
    sigma = torch.relu(sigma)
    if x_dir is None:
        return None, sigma
    h = torch.relu(lin("fc10", torch.cat([h[:, :hidden], x_dir], dim=1)))
    return torch.sigmoid(lin("fc11", h)), sigma


class NeRFOracle:
    def __init__(self, cfg, frame: HorizontalFrame | None, geo=None):
        """`geo` = (scale, offset (3,) float64, ray_origin_height): what pipeline.py:30-58 captures from
        the dataset; needed when there is no frame (no point preprocessor) or with `include_height`."""
        self.cfg, self.frame, self.geo = cfg, frame, geo
        self.lx, self.ld = cfg["encoder"]["L_x"], cfg["encoder"]["L_d"]
        self.hidden = cfg["mlp_hidden_dim"]
        self.height = bool(cfg.get("include_height", False))
        n_pos = 4 if self.height else 3                                   # models/nerf.py:126-133
        pos_ch = sum(self.lx) * 2 if isinstance(self.lx, list) else self.lx * 2 * n_pos
        self.shapes = {
            "coarse": mlp_shapes(pos_ch, self.ld * 6, cfg["num_bands"], 1, self.hidden),
            "fine": mlp_shapes(pos_ch, self.ld * 6, cfg["num_bands"], cfg["num_bands"], self.hidden),
        }

    def init_params(self, seed=0):
        gen = torch.Generator().manual_seed(seed)
        return {m: {k: v.requires_grad_() for k, v in init_mlp(self.shapes[m], gen).items()} for m in ("coarse", "fine")}

    def _encode_pos(self, pts):
        if isinstance(self.lx, list):
            return pe_per_axis(pts, self.lx)
        return pe_interleaved(pts, self.lx)

    def _stage(self, mode, batch, sd, z, noise):
        """pipelines/nerf.py:122-167 after the sampling step."""
        b, n = z.shape
        pts = sampling.points_on_rays(batch["origin"], batch["dir"], z)
        if self.frame is not None:
            pts = preprocess_horizontal(pts, self.frame)
        if self.height:  # pipelines/nerf.py:127-128
            scale, offset, h0 = self.geo
            pts = sampling.append_heights(pts, h0, scale, offset)
        x_pos = self._encode_pos(pts).view(b * n, -1)
        dirs = batch["dir"][:, None].repeat(1, n, 1)
        x_dir = pe_interleaved(dirs, self.ld).view(b * n, -1)
        color, sigma = mlp_forward(sd, x_pos, x_dir, self.hidden, noise)
        color = torch.exp(torch.clamp(color.view(b, n, -1), max=11))
        sigma = torch.relu(sigma.view(b, n, -1))
        scale = self.frame.scale if self.frame is not None else self.geo[0]
        c, _, w = rendering.composite(z * (scale / 1000), color, sigma)
        return {"color": color, "sigma": sigma, "color_map": c, "weights": w, "z_vals": z}

    def forward(self, batch, params, u_c, u_f, noise_c=None, noise_f=None):
        """pipelines/nerf.py:169-188."""
        nc = self.cfg["sampler"]["N_c"]
        z_c = sampling.stratified_z(batch["len"], nc, u_c)
        rc = self._stage("coarse", batch, params["coarse"], z_c, noise_c)
        z_f, inds = sampling.inverse_cdf_z(rc["weights"], rc["z_vals"], u_f)
        rf = self._stage("fine", batch, params["fine"], z_f, noise_f)
        out = {f"{k}_coarse": v for k, v in rc.items()}
        out.update({f"{k}_fine": v for k, v in rf.items()})
        out["inds_fine"] = inds
        return out

    def loss(self, batch, res):
        """pipelines/nerf.py:219-240."""
        pc = rendering.band_select(res["color_map_coarse"], batch["irgb_idx"])
        pf = rendering.band_select(res["color_map_fine"], batch["irgb_idx"])
        mse = torch.nn.functional.mse_loss
        return mse(pc, batch["rad"]) + mse(pf, batch["rad"])

    def extract(self, pts, params):
        """pipelines/nerf.py:190-217."""
        if self.frame is not None:
            pts = preprocess_horizontal(pts[None], self.frame)[0]
        if self.height:  # pipelines/nerf.py:205-208
            scale, offset, h0 = self.geo
            pts = sampling.append_heights(pts[None], h0, scale, offset)[0]
        x_pos = self._encode_pos(pts).view(pts.shape[0], -1).float()
        _, sigma = mlp_forward(params["fine"], x_pos, None, self.hidden)
        return torch.clip(sigma, min=0)
