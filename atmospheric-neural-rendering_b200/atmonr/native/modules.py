"""Drop-in replacements for the two tiny-cuda-nn classes the reference instantiates:

    tcnn.Encoding(n_input_dims, cfg)         -> atmonr.native.modules.Encoding
    tcnn.Network(n_input_dims, n_out, cfg)   -> atmonr.native.modules.Network

(reference: src/atmonr/pipelines/instant_ngp.py:60-85). Same constructor arguments, same
`.n_output_dims`, same nn.Module protocol with ONE flat float32 parameter named `params`
(so `state_dict()` is `{"params": tensor}` like tcnn's and checkpoints interchange).

Deliberate difference: outputs are float32 tensors holding the result of the fp16-operand /
fp32-accumulate arithmetic (tcnn returns float16). Gradients therefore flow in float32 and do
not need tcnn's loss scaling.
"""

from __future__ import annotations

import math

import torch
from torch import nn

from atmonr.native import lib as L
from atmonr.native import ops


def default_device() -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def shadow_of(param: torch.Tensor) -> torch.Tensor:
    """fp16 shadow of a float32 parameter, refreshed whenever torch's version counter says the
    parameter changed. The fused optimizer writes parameter and shadow in one pass through raw
    pointers (no version bump), so the pair stays consistent without a separate cast."""
    sh = getattr(param, "_atmonr_shadow", None)
    if sh is None or getattr(param, "_atmonr_shadow_version", None) != param._version or sh.device != param.device:
        sh = param.detach().to(torch.float16).contiguous()
        param._atmonr_shadow = sh
        param._atmonr_shadow_version = param._version
    return sh


def _sh2(x: torch.Tensor) -> torch.Tensor:
    """tcnn SphericalHarmonics degree 2 (4 outputs) on 2x-1; fp16-rounded like the kernels."""
    v = x * 2 - 1
    c0, c1 = 0.28209479177387814, 0.48860251190291987
    out = torch.stack([torch.full_like(v[:, 0], c0), -c1 * v[:, 1], c1 * v[:, 2], -c1 * v[:, 0]], dim=1)
    return out.half().float()


class Encoding(nn.Module):
    """HashGrid, or Composite of {HashGrid, SphericalHarmonics(degree 2), Identity}."""

    def __init__(self, n_input_dims: int, encoding_config: dict, seed: int = 1337, device=None):
        super().__init__()
        self.n_input_dims = n_input_dims
        self.encoding_config = encoding_config
        self.seed = seed
        device = device or default_device()
        cfg = encoding_config
        nested = cfg["nested"] if cfg["otype"] == "Composite" else [dict(cfg, n_dims_to_encode=n_input_dims)]
        self.parts = []  # (kind, lo, hi, grid | None)
        lo = 0
        for sub in nested:
            n = int(sub.get("n_dims_to_encode", n_input_dims - lo))
            kind = sub["otype"]
            grid = None
            if kind == "HashGrid":
                if n not in (2, 3, 4):  # 4: positions + height (`include_height`)
                    raise NotImplementedError("HashGrid over 2, 3 or 4 dims")
                grid = L.grid_layout(n, sub)
            elif kind == "SphericalHarmonics":
                if int(sub["degree"]) != 2 or n != 3:
                    raise NotImplementedError("SphericalHarmonics: degree 2 over 3 dims only")
            elif kind != "Identity":
                raise NotImplementedError(f"encoding otype {kind}")
            self.parts.append((kind, lo, lo + n, grid))
            lo += n
        if lo != n_input_dims:
            raise ValueError("nested encodings do not cover the input dims")
        grids = [g for k, _, _, g in self.parts if k == "HashGrid"]
        if len(grids) > 1:
            raise NotImplementedError("at most one HashGrid per encoding")
        self.grid = grids[0] if grids else None
        self.n_output_dims = sum(
            2 * g.n_levels if k == "HashGrid" else (4 if k == "SphericalHarmonics" else hi - lo_)
            for k, lo_, hi, g in self.parts
        )
        n_params = 2 * self.grid.n_entries if self.grid else 0
        gen = torch.Generator().manual_seed(seed)
        init = (torch.rand(n_params, generator=gen) * 2 - 1) * 1e-4  # tcnn: U(-1e-4, 1e-4)
        self.params = nn.Parameter(init.to(device))

    def table_f16(self) -> torch.Tensor:
        return shadow_of(self.params)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x.float()
        outs = []
        for kind, lo, hi, grid in self.parts:
            xs = x[:, lo:hi]
            if kind == "HashGrid":
                outs.append(ops.HashGridFn.apply(xs.contiguous(), self.params, self.table_f16(), grid))
            elif kind == "SphericalHarmonics":
                outs.append(_sh2(xs))
            else:
                outs.append(xs.half().float())
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)


class Network(nn.Module):
    """Bias-free fully fused MLP: ReLU hidden layers of width 32, no output activation."""

    def __init__(self, n_input_dims: int, n_output_dims: int, network_config: dict, seed: int = 1337, device=None):
        super().__init__()
        self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims
        self.network_config = network_config
        self.seed = seed
        self.shape = L.mlp_shape(n_input_dims, n_output_dims, network_config)
        if self.shape.width != 32 or self.shape.out_pad != 16:
            raise NotImplementedError("n_neurons must be 32 and n_output_dims <= 16")
        device = device or default_device()
        gen = torch.Generator().manual_seed(seed)
        s = self.shape
        dims = [(s.width, s.in_pad)] + [(s.width, s.width)] * (s.n_hidden - 1) + [(s.out_pad, s.width)]
        chunks = [(torch.rand(o * i, generator=gen) * 2 - 1) * math.sqrt(6.0 / (o + i)) for o, i in dims]
        self.params = nn.Parameter(torch.cat(chunks).to(device))  # Xavier-uniform per matrix

    def weights_f16(self) -> torch.Tensor:
        return shadow_of(self.params)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.MlpFn.apply(x.float(), self.params, self.weights_f16(), self.shape)
