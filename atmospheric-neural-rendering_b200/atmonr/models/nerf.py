"""AtmoNeRF MLP (reference: src/atmonr/models/nerf.py).

Eleven biased linear layers, width `hidden_dim`, a skip connection that re-injects the encoded
position at layer 6, a density head on layer 9 (with unit Gaussian noise while training) and a
direction-conditioned colour head. The layers are torch.nn.Linear modules, so parameter names
(`fc1`..`fc11`) match the reference and checkpoints interchange. On the device forward, input
gradient and weight gradient of every layer run on tcgen05 (csrc/linear_tc.cu: float32 operands
split into bfloat16 terms, partial products accumulated in float32 in TMEM; validated on a B200 in
round 2, tests/test_zz_gpu_linear_tc.py), the whole network as one autograd node with a hand-written
backward chain (atmonr.native.nerf_mlp). `DENSE_IMPL = "library"` switches the layers to
torch's float32 GEMMs: used by the tests and by scripts/bench_nerf.py as the cross-check, never by
the pipelines.
"""

from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


DENSE_IMPL = "tc"   # "tc" (product path) | "library" (cross-check in tests / bench only)


class AtmoNeRF(nn.Module):
    def __init__(self, pos_channels: int, dir_channels: int, out_channels: int, volume_channels: int,
                 hidden_dim: int = 256) -> None:
        super().__init__()
        self.pos_channels, self.dir_channels = pos_channels, dir_channels
        self.out_channels, self.volume_channels, self.hidden_dim = out_channels, volume_channels, hidden_dim
        h = hidden_dim
        widths = [(pos_channels, h), (h, h), (h, h), (h, h), (h, h), (h + pos_channels, h), (h, h), (h, h),
                  (h, h + volume_channels), (h + dir_channels, h // 2), (h // 2, out_channels)]
        for k, (fan_in, fan_out) in enumerate(widths, start=1):
            layer = nn.Linear(fan_in, fan_out)
            nn.init.kaiming_normal_(layer.weight, mode="fan_out")
            setattr(self, f"fc{k}", layer)

    def _layer(self, k: int, x: torch.Tensor, relu: bool, x2: torch.Tensor | None = None) -> torch.Tensor:
        """relu?(fc_k(cat([x, x2]))); on the tensor-core path the concatenation is never materialised."""
        fc = getattr(self, f"fc{k}")
        if x.is_cuda and DENSE_IMPL == "tc":
            from atmonr.native import ops
            return ops.linear_tc(x, fc.weight, fc.bias, relu, x2=x2)
        if x2 is not None:
            x = torch.cat([x, x2], dim=1)
        return F.relu(fc(x)) if relu else fc(x)

    def forward_pos_only(self, x_pos: torch.Tensor):
        """Trunk up to the density head. models/nerf.py:48-73."""
        x = x_pos
        for k in range(1, 6):
            x = self._layer(k, x, True)
        x = self._layer(6, x, True, x2=x_pos)
        x = self._layer(7, x, True)
        x = self._layer(8, x, True)
        x = self._layer(9, x, False)
        sigma = x[:, self.hidden_dim:]
        if self.training:
            sigma = sigma + torch.randn(sigma.shape, device=sigma.device)
        return x, F.relu(sigma)

    def forward(self, x: torch.Tensor):
        """models/nerf.py:75-93 -> (colour in (0,1), density >= 0)."""
        if x.is_cuda and DENSE_IMPL == "tc":
            # the whole MLP as one autograd node over the tensor-core products (atmonr.native.nerf_mlp)
            from atmonr.native.nerf_mlp import NerfMlpFn
            noise = torch.randn((x.shape[0], self.volume_channels), device=x.device) if self.training else None
            params = [t for k in range(1, 12) for t in (getattr(self, f"fc{k}").weight, getattr(self, f"fc{k}").bias)]
            rgb, sigma = NerfMlpFn.apply(x, noise, self.pos_channels, self.hidden_dim, *params)
            return torch.sigmoid(rgb), F.relu(sigma)
        x_pos, d = x[:, : self.pos_channels], x[:, self.pos_channels:]
        feat, sigma = self.forward_pos_only(x_pos)
        hid = self._layer(10, feat[:, : self.hidden_dim], True, x2=d)
        return torch.sigmoid(self._layer(11, hid, False)), sigma


def get_model(hidden_dim: int, N_lambda: int, L_x, L_d: int, include_height: bool):
    """models/nerf.py:96-144 -> (coarse, fine): the coarse net has one density, the fine net one
    density per band."""
    if isinstance(L_x, int):
        pos_channels = L_x * (8 if include_height else 6)
    else:
        assert len(L_x) == (4 if include_height else 3)
        pos_channels = 2 * sum(L_x)
    common = dict(pos_channels=pos_channels, dir_channels=6 * L_d, out_channels=N_lambda, hidden_dim=hidden_dim)
    return AtmoNeRF(volume_channels=1, **common), AtmoNeRF(volume_channels=N_lambda, **common)
