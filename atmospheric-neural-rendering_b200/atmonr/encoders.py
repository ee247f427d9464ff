"""NeRF positional encoding (reference: src/atmonr/encoders.py:4-28)."""

from __future__ import annotations

import torch

from atmonr.native import ops


class _PositionalEncodingFn(torch.autograd.Function):
    """sin/cos features of 2^l * pi * p. Backward is analytic: d sin = cos * f, d cos = -sin * f."""

    @staticmethod
    def forward(ctx, flat, freqs, interleaved):
        out = ops.positional_encoding(flat, freqs, interleaved)
        ctx.save_for_backward(flat, out)
        ctx.freqs, ctx.interleaved = freqs, interleaved
        return out

    @staticmethod
    def backward(ctx, g):
        flat, out = ctx.saved_tensors
        grads, col = [], 0
        for axis, n in enumerate(ctx.freqs):
            f = (2.0 ** torch.arange(n, device=flat.device, dtype=flat.dtype)) * torch.pi
            blk, gb = out[:, col:col + 2 * n], g[:, col:col + 2 * n]
            if ctx.interleaved:
                s, c, gs, gc = blk[:, 0::2], blk[:, 1::2], gb[:, 0::2], gb[:, 1::2]
            else:
                s, c, gs, gc = blk[:, :n], blk[:, n:], gb[:, :n], gb[:, n:]
            grads.append(((gs * c - gc * s) * f).sum(dim=1))
            col += 2 * n
        return torch.stack(grads, dim=1), None, None


def positional_encoding(pts: torch.Tensor, L: int | list[int]) -> torch.Tensor:
    """Same two layouts as the reference:
    int L   -> shape (M, C, 2L), per axis [sin, cos] interleaved per frequency;
    list L  -> shape (..., 2*sum(L)), per axis [sin x L_i | cos x L_i]."""
    c = pts.shape[-1]
    flat = pts.reshape(-1, c)
    if isinstance(L, int):
        out = _PositionalEncodingFn.apply(flat, tuple([L] * c), True)
        return out.view(flat.shape[0], c, 2 * L)
    if isinstance(L, list):
        out = _PositionalEncodingFn.apply(flat, tuple(L), False)
        return out.view(*pts.shape[:-1], 2 * sum(L))
    raise TypeError("L must be an int or a list of ints")
