"""Oracle: restatement of the tiny-cuda-nn modules the reference instantiates.

TEST INFRASTRUCTURE (see oracle/__init__.py).  **PARITY UNPINNED**: tiny-cuda-nn
(NVlabs/tiny-cuda-nn, torch binding `tinycudann`, installed un-pinned from git HEAD by the
reference, README.md:19-22) is not vendored, not installed and not installable here; the
reference holds no test or golden vector for it.  Everything below marked [upstream-recalled]
restates the published algorithm (Mueller et al., "Instant Neural Graphics Primitives", 2022)
and the upstream source as recalled in SURVEY.md section 8c.  Call sites that anchor it:
src/atmonr/pipelines/instant_ngp.py:60-85 (construction), :163-174 (training forward),
:236-237 (extract), configs/instant_ngp.json:18-81 (module configs).

Two numeric modes:
  * ``fp16=False``: everything in float32 -- the ground truth of the tolerance tests.
  * ``fp16=True`` : emulates the rounding points of the native sm_100a path (table entries,
    MLP weights, encoder outputs and hidden activations rounded to fp16; fp32 accumulate in the
    MLPs; the hash-grid interpolation accumulates in fp16 with one fused multiply-add per corner,
    as tiny-cuda-nn does),
    with straight-through gradients, so the CUDA kernels can be checked tightly.
"""

from __future__ import annotations

import math

import numpy as np
import torch

PRIMES = (1, 2654435761, 805459861, 3674653429)  # [upstream-recalled] coherent prime hash


def _ste_half(x: torch.Tensor) -> torch.Tensor:
    """Round to fp16 with a straight-through gradient."""
    return x + (x.half().float() - x).detach()


# ------------------------------------------------------------------------------------------
# multiresolution hash grid
# ------------------------------------------------------------------------------------------
def grid_levels(n_dims, n_levels, log2_hashmap_size, base_resolution, per_level_scale):
    """[upstream-recalled] level table of tcnn's GridEncoding (grid.h).

    scale_l = exp2(l * log2(per_level_scale)) * base - 1, evaluated in float64 on the float32
              per_level_scale and rounded once to float32 (tcnn evaluates it in float32 with
              exp2f/log2f, whose host and device versions disagree in the last bits; the
              once-rounded value is within a few ulps and reproducible everywhere)
    res_l   = ceilf(scale_l) + 1
    size_l  = min(next_multiple(res_l ** D, 8), 2 ** log2T)  (entries, each F features)
    a level is 'hashed' iff res_l ** D (stride after D dims) exceeds size_l.
    """
    f32 = np.float32
    log2_pls = np.log2(np.float64(f32(per_level_scale)))
    scale = np.zeros(n_levels, dtype=f32)
    res = np.zeros(n_levels, dtype=np.uint32)
    size = np.zeros(n_levels, dtype=np.uint32)
    offset = np.zeros(n_levels + 1, dtype=np.uint32)
    cap = 1 << log2_hashmap_size
    max_params = (2**32 - 1) // 2
    for lvl in range(n_levels):
        s = f32(np.exp2(np.float64(lvl) * log2_pls) * np.float64(base_resolution) - 1.0)
        scale[lvl] = s
        r = int(np.ceil(s)) + 1
        res[lvl] = r
        dense = r**n_dims
        n = max_params if float(r) ** n_dims > float(max_params) else dense
        n = (n + 7) // 8 * 8
        n = min(n, cap)
        size[lvl] = n
        offset[lvl + 1] = offset[lvl] + n
    return {"scale": scale, "res": res, "size": size, "offset": offset}


def grid_corner_indices(x01: torch.Tensor, levels: dict, level: int):
    """[upstream-recalled] kernel_grid / grid_index for one level.

    x01: (M, D) float32.  Returns (idx (M, 2^D) int64 *within the level*, w (M, 2^D) f32).
    pos = fmaf(scale, x, 0.5f) is evaluated in float64 and rounded once to float32 (exact
    product; bit-identical to the fused op up to double-rounding cases of measure ~2^-29).
    """
    m, d = x01.shape
    scale = float(levels["scale"][level])
    res = int(levels["res"][level])
    size = int(levels["size"][level])
    pos = (x01.double() * scale + 0.5).float()
    cell = torch.floor(pos)
    frac = pos - cell
    cell = cell.to(torch.int64) & 0xFFFFFFFF  # (uint32)(int) cast
    idx_out, w_out = [], []
    for corner in range(1 << d):
        w = torch.ones(m, dtype=torch.float32)
        g = []
        for k in range(d):
            if corner & (1 << k):
                w = w * frac[:, k]
                g.append((cell[:, k] + 1) & 0xFFFFFFFF)
            else:
                w = w * (1 - frac[:, k])
                g.append(cell[:, k])
        stride, index = 1, torch.zeros(m, dtype=torch.int64)
        k = 0
        while k < d and stride <= size:
            index = (index + g[k] * stride) & 0xFFFFFFFF
            stride *= res
            k += 1
        if size < stride:  # hashed level
            index = torch.zeros(m, dtype=torch.int64)
            for k in range(d):
                index = index ^ ((g[k] * PRIMES[k]) & 0xFFFFFFFF)
        idx_out.append(index % size)
        w_out.append(w)
    return torch.stack(idx_out, 1), torch.stack(w_out, 1)


class HashGrid:
    """tcnn.Encoding(n, {"otype": "HashGrid", ...}) -- instant_ngp.py:60-63,78-80."""

    def __init__(self, n_dims, cfg, n_features=None):
        self.n_dims = n_dims
        self.n_levels = int(cfg["n_levels"])
        self.n_feat = int(cfg.get("n_features_per_level", 2))
        self.levels = grid_levels(
            n_dims, self.n_levels, int(cfg["log2_hashmap_size"]),
            int(cfg["base_resolution"]), float(cfg["per_level_scale"]),
        )
        self.n_entries = int(self.levels["offset"][-1])
        self.n_params = self.n_entries * self.n_feat
        self.n_output_dims = self.n_levels * self.n_feat

    def init_params(self, gen: torch.Generator) -> torch.Tensor:
        """[upstream-recalled] U(-1e-4, 1e-4)."""
        return (torch.rand(self.n_params, generator=gen) * 2 - 1) * 1e-4

    def forward(self, x01, params, fp16=False):
        tab = params.view(self.n_entries, self.n_feat)
        if fp16:
            tab = _ste_half(tab)
        outs = []
        for lvl in range(self.n_levels):
            idx, w = grid_corner_indices(x01.detach(), self.levels, lvl)
            base = int(self.levels["offset"][lvl])
            acc = torch.zeros(x01.shape[0], self.n_feat, dtype=torch.float32)
            for c in range(idx.shape[1]):
                acc = acc + w[:, c, None] * tab[base + idx[:, c]]
            if fp16:
                # [upstream-recalled] tcnn grid.h: result = fma((half)weight, entry, result), all in
                # fp16, corners in index order. One rounding per corner: the product of two fp16
                # values and its sum with an fp16 accumulator are exact in float64.
                # (numpy rounds float64 -> float16 once; torch goes through float32, a double rounding)
                chain = np.zeros((x01.shape[0], self.n_feat), np.float16)
                with torch.no_grad():
                    w16 = w.numpy().astype(np.float16).astype(np.float64)
                    t64 = tab.detach().numpy().astype(np.float64)
                    for c in range(idx.shape[1]):
                        prod = w16[:, c, None] * t64[base + idx[:, c].numpy()]
                        chain = (chain.astype(np.float64) + prod).astype(np.float16)
                chain = torch.from_numpy(chain.astype(np.float32))
                acc = chain + (acc - acc.detach())   # value: the fp16 chain, exactly; gradient: straight through
            outs.append(acc)
        return torch.cat(outs, dim=1)

    def all_indices(self, x01):
        """Global entry index of every (sample, level, corner): (M, L, 2^D) int64."""
        out = []
        for lvl in range(self.n_levels):
            idx, _ = grid_corner_indices(x01, self.levels, lvl)
            out.append(idx + int(self.levels["offset"][lvl]))
        return torch.stack(out, 1)


# ------------------------------------------------------------------------------------------
# the other encodings
# ------------------------------------------------------------------------------------------
def sh_degree2(x):
    """[upstream-recalled] SphericalHarmonics degree 2 = 4 outputs; tcnn maps its input from
    [0,1] to [-1,1] first.  The reference feeds raw unit vectors (instant_ngp.py:157,165-169)
    -- that quirk is kept by the caller, not corrected here."""
    v = x * 2 - 1
    c0, c1 = 0.28209479177387814, 0.48860251190291987
    return torch.stack(
        [torch.full_like(v[:, 0], c0), -c1 * v[:, 1], c1 * v[:, 2], -c1 * v[:, 0]], dim=1
    )


class Composite:
    """tcnn Composite encoding restricted to the nested kinds the reference configures
    (configs/instant_ngp.json:33-44,52-70): HashGrid, SphericalHarmonics(degree 2), Identity.
    A nested entry without n_dims_to_encode takes the remaining dims [upstream-recalled]."""

    def __init__(self, n_dims, cfg):
        self.parts = []  # (kind, lo, hi, obj)
        lo = 0
        for sub in cfg["nested"]:
            n = int(sub.get("n_dims_to_encode", n_dims - lo))
            kind = sub["otype"]
            obj = HashGrid(n, sub) if kind == "HashGrid" else None
            if kind == "SphericalHarmonics":
                assert int(sub["degree"]) == 2 and n == 3
            self.parts.append((kind, lo, lo + n, obj))
            lo += n
        assert lo == n_dims
        self.n_output_dims = sum(
            o.n_output_dims if k == "HashGrid" else (4 if k == "SphericalHarmonics" else hi - lo_)
            for k, lo_, hi, o in self.parts
        )
        self.n_params = sum(o.n_params for k, _, _, o in self.parts if k == "HashGrid")

    def grid(self):
        for k, _, _, o in self.parts:
            if k == "HashGrid":
                return o
        return None

    def init_params(self, gen):
        g = self.grid()
        return g.init_params(gen) if g else torch.zeros(0)

    def forward(self, x, params, fp16=False):
        outs = []
        for kind, lo, hi, obj in self.parts:
            xs = x[:, lo:hi]
            if kind == "HashGrid":
                outs.append(obj.forward(xs, params, fp16))
            elif kind == "SphericalHarmonics":
                y = sh_degree2(xs)
                outs.append(_ste_half(y) if fp16 else y)
            elif kind == "Identity":
                outs.append(_ste_half(xs) if fp16 else xs)
            else:
                raise NotImplementedError(kind)
        return torch.cat(outs, dim=1)


def make_encoding(n_dims, cfg):
    """tcnn.Encoding(n_dims, cfg)."""
    if cfg["otype"] == "HashGrid":
        return HashGrid(n_dims, cfg)
    if cfg["otype"] == "Composite":
        return Composite(n_dims, cfg)
    raise NotImplementedError(cfg["otype"])


# ------------------------------------------------------------------------------------------
# FullyFusedMLP
# ------------------------------------------------------------------------------------------
class Network:
    """tcnn.Network(n_in, n_out, {"otype": "FullyFusedMLP", ...}) -- instant_ngp.py:64-85.

    [upstream-recalled] no biases; ReLU hidden activation, no output activation; the input
    is padded to a multiple of 16 WITH 1.0 (so the weights of the padded columns act as a
    bias), the output to a multiple of 16; weights are row-major [out][in], concatenated
    first to last in one flat fp32 `params` vector; Xavier-uniform initialisation."""

    def __init__(self, n_in, n_out, cfg):
        assert cfg["activation"] == "ReLU" and cfg["output_activation"] == "None"
        self.n_in, self.n_out = n_in, n_out
        self.width = int(cfg["n_neurons"])
        self.n_hidden = int(cfg["n_hidden_layers"])
        self.in_pad = (n_in + 15) // 16 * 16
        self.out_pad = (n_out + 15) // 16 * 16
        self.shapes = [(self.width, self.in_pad)]
        self.shapes += [(self.width, self.width)] * (self.n_hidden - 1)
        self.shapes += [(self.out_pad, self.width)]
        self.n_params = sum(o * i for o, i in self.shapes)
        self.n_output_dims = n_out

    def init_params(self, gen):
        chunks = []
        for o, i in self.shapes:
            bound = math.sqrt(6.0 / (o + i))
            chunks.append((torch.rand(o * i, generator=gen) * 2 - 1) * bound)
        return torch.cat(chunks)

    def matrices(self, params):
        mats, at = [], 0
        for o, i in self.shapes:
            mats.append(params[at : at + o * i].view(o, i))
            at += o * i
        return mats

    def forward(self, x, params, fp16=False):
        """x: (M, n_in).  Returns (M, n_out) float32.  With fp16=True the input and every
        hidden activation are rounded to fp16, weights are fp16, accumulation is fp32 and
        the OUTPUT stays fp32 (the native path reads it straight from the accumulator)."""
        m = x.shape[0]
        if self.in_pad > self.n_in:
            x = torch.cat([x, torch.ones(m, self.in_pad - self.n_in, dtype=x.dtype)], dim=1)
        mats = self.matrices(params)
        h = _ste_half(x) if fp16 else x
        for k, w in enumerate(mats):
            wk = _ste_half(w) if fp16 else w
            h = h @ wk.t()
            if k < len(mats) - 1:
                h = torch.relu(h)
                if fp16:
                    h = _ste_half(h)
        return h[:, : self.n_out]
