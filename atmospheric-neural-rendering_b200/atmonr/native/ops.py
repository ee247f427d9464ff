"""Operator layer over libatmonr_b200: tensor-in / tensor-out functions and the
torch.autograd.Function wrappers the pipelines are built from.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); every arithmetic
step on the hot path is a kernel of the shared library. No function in this module has a
CPU implementation: CPU tensors raise.
"""

from __future__ import annotations

import ctypes as C
import os

import torch

from atmonr.native import lib as L

_f32 = torch.float32

LOSS_KINDS = {"dark": 0, "hdr": 1, "l1": 2, "l1_plus_hdr": 3, "mse": 4, "mse_plus_hdr": 5}


def _c(t: torch.Tensor, dtype=None) -> torch.Tensor:
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def linspace_bins(n_bins: int, device) -> torch.Tensor:
    """linspace(0, 1, n+1)[:-1], computed by torch so the kernel uses the reference's exact
    bin edges (samplers.py:34)."""
    return torch.linspace(0, 1, n_bins + 1, device=device)[:-1].contiguous()


# ------------------------------------------------------------------------------------------
# samplers / preprocessor
# ------------------------------------------------------------------------------------------
def sample_uniform(origin, direction, length, n_bins, u=None, random=True, seed=0, ray_index_base=0):
    """samplers.py:8-47. u given -> mode 1; random and no u -> in-kernel Philox; else mid-points."""
    origin, direction, length = _c(origin, _f32), _c(direction, _f32), _c(length, _f32)
    b = origin.shape[0]
    mode = 1 if u is not None else (2 if random else 0)
    if u is not None:
        u = _c(u, _f32)
        assert u.shape == (b, n_bins)
    pts = torch.empty((b, n_bins, 3), device=origin.device, dtype=_f32)
    z = torch.empty((b, n_bins), device=origin.device, dtype=_f32)
    bins = linspace_bins(n_bins, origin.device)
    L.call("atmonr_sample_uniform", L.ptr(origin), L.ptr(direction), L.ptr(length), L.ptr(u), L.ptr(bins),
           b, n_bins, mode, seed, ray_index_base, L.ptr(pts), L.ptr(z), L.stream())
    return pts, z


def preprocess_horizontal(frame: L.FrameT, pts: torch.Tensor) -> torch.Tensor:
    """harp2.py:372-386 on float32 or float64 points of shape (..., 3)."""
    if pts.dtype not in (torch.float32, torch.float64):
        pts = pts.float()
    p = pts.contiguous()
    out = torch.empty_like(p)
    L.call("atmonr_preprocess_horizontal", C.byref(frame), L.ptr(p), L.ptr(out), p.numel() // 3,
           int(p.dtype == torch.float64), L.stream())
    return out


def ngp_sample_points(frame, origin, direction, length, n, alt_compress, u=None, random=True, seed=0,
                      ray_index_base=0, bins=None, out=None, height=None):
    """Fused instant_ngp.py:139-160 -> (x01 (B*n,3), z (B,n)); `out` = preallocated (x01, z).
    height = (scale, offset (3 floats), ray_origin_height): `include_height`, x01 gets a fourth column
    (atmonr_ngp_sample_points_height)."""
    origin, direction, length = _c(origin, _f32), _c(direction, _f32), _c(length, _f32)
    b = origin.shape[0]
    d = 4 if height is not None else 3
    mode = 1 if u is not None else (2 if random else 0)
    if u is not None:
        u = _c(u, _f32)
    if out is not None:
        x01, z = out
        assert x01.shape == (b * n, d) and z.shape == (b, n) and x01.dtype == _f32 and z.dtype == _f32
    else:
        x01 = torch.empty((b * n, d), device=origin.device, dtype=_f32)
        z = torch.empty((b, n), device=origin.device, dtype=_f32)
    if bins is None:
        bins = linspace_bins(n, origin.device)
    if height is not None:
        scale, offset, h0 = height
        off = (C.c_double * 3)(*[float(v) for v in offset])
        L.call("atmonr_ngp_sample_points_height", C.byref(frame), L.ptr(origin), L.ptr(direction), L.ptr(length), L.ptr(u),
               L.ptr(bins), b, n, mode, seed, ray_index_base, float(alt_compress), float(scale), off, float(h0),
               L.ptr(x01), L.ptr(z), L.stream())
    else:
        L.call("atmonr_ngp_sample_points", C.byref(frame), L.ptr(origin), L.ptr(direction), L.ptr(length), L.ptr(u),
               L.ptr(bins), b, n, mode, seed, ray_index_base, float(alt_compress), L.ptr(x01), L.ptr(z), L.stream())
    return x01, z


# ------------------------------------------------------------------------------------------
# ray table and batch gather (dataset side of the path)
# ------------------------------------------------------------------------------------------
def get_rays(lat, lon, alt, thetav, phiv, ray_origin_height, tol: float = 10.0, max_iters: int = 20):
    """wgs_84.py:223-290 for one chunk of pixels: (P,A) float32 inputs -> origins (P*A,3),
    directions (P*A,3), lengths (P*A,), all float32. Synchronises the current stream."""
    lat, lon, alt, thetav, phiv = (_c(t, _f32) for t in (lat, lon, alt, thetav, phiv))
    if not (lat.shape == lon.shape == alt.shape == thetav.shape == phiv.shape):
        raise ValueError("lat, lon, alt, thetav, phiv must share a shape")
    n = lat.numel()
    dev = lat.device
    origin = torch.empty((n, 3), device=dev, dtype=_f32)
    direction = torch.empty((n, 3), device=dev, dtype=_f32)
    length = torch.empty((n,), device=dev, dtype=_f32)
    work = torch.empty((2 * n + 1,), device=dev, dtype=torch.float64)
    iters = C.c_int(0)
    L.call("atmonr_get_rays", L.ptr(lat), L.ptr(lon), L.ptr(alt), L.ptr(thetav), L.ptr(phiv), n,
           float(ray_origin_height), float(tol), int(max_iters), L.ptr(origin), L.ptr(direction), L.ptr(length),
           L.ptr(work), C.byref(iters), L.stream())
    get_rays.last_iters = iters.value
    return origin, direction, length


def filter_rays(ray_origin, ray_dir, ray_rad):
    """wgs_84.py:293-313: torch.bool mask of the rays without a NaN in origin, direction or radiance."""
    ray_origin, ray_dir, ray_rad = (_c(t, _f32) for t in (ray_origin, ray_dir, ray_rad))
    n = ray_rad.numel()
    if ray_origin.shape != (n, 3) or ray_dir.shape != (n, 3):
        raise ValueError("ray_origin and ray_dir must be (n, 3) for n radiances")
    valid = torch.empty((n,), device=ray_rad.device, dtype=torch.bool)
    L.call("atmonr_filter_rays", L.ptr(ray_origin), L.ptr(ray_dir), L.ptr(ray_rad), n, L.ptr(valid), L.stream())
    return valid


RAY_EXTENT_WORK_BYTES = 37888   # ATMONR_RAY_EXTENT_WORK_BYTES of include/atmonr_b200.h


def normalize_rays(ray_origin, ray_dir, ray_len):
    """wgs_84.py:316-339: (origins normalised into [-1, 1]^3 float32, scale as a Python float, offset
    as a float64[3] device tensor). The bounding box and the normalisation are kernels; the two lines
    between them (:336-337, six numbers) are the reference's own float64 expressions."""
    ray_origin, ray_dir, ray_len = (_c(t, _f32) for t in (ray_origin, ray_dir, ray_len))
    n = ray_len.numel()
    if ray_origin.shape != (n, 3) or ray_dir.shape != (n, 3):
        raise ValueError("ray_origin and ray_dir must be (n, 3) for n lengths")
    dev = ray_origin.device
    hi_lo = torch.empty((6,), device=dev, dtype=_f32)
    work = torch.empty((RAY_EXTENT_WORK_BYTES // 4,), device=dev, dtype=_f32)
    L.call("atmonr_ray_extent", L.ptr(ray_origin), L.ptr(ray_dir), L.ptr(ray_len), n, L.ptr(hi_lo), L.ptr(work),
           L.stream())
    hi, lo = hi_lo[:3].double(), hi_lo[3:].double()
    scale = ((hi - lo).max() / 2).item()
    offset = ((hi + lo) / 2).contiguous()
    out = torch.empty_like(ray_origin)
    L.call("atmonr_normalize_origins", L.ptr(ray_origin), n, L.ptr(offset), float(scale), L.ptr(out), L.stream())
    return out, scale, offset


def gather_batch(tables: dict, index: torch.Tensor) -> dict:
    """harp2.py:392-420: the per-ray tables {origin, dir, alt, rad, len, idx, irgb_idx} gathered at
    `index` (int64, negative values count from the end) in one launch. An out-of-range index raises
    IndexError when `gather_batch.check` is set (costs a device->host read per batch); the kernel never
    reads out of bounds either way."""
    index = _c(index, torch.int64)
    if index.dim() != 1:
        raise ValueError("batch index must be 1-D")
    b, r = index.shape[0], tables["origin"].shape[0]
    dev = index.device
    want = {"origin": (_f32, (r, 3)), "dir": (_f32, (r, 3)), "alt": (_f32, (r,)), "rad": (_f32, (r,)),
            "len": (_f32, (r,)), "idx": (torch.int32, (r,)), "irgb_idx": (torch.int64, (r,))}
    for k, (dt, shp) in want.items():
        t = tables[k]
        if t.dtype != dt or tuple(t.shape) != shp or not t.is_contiguous():
            raise L.NativeLibraryError(f"gather_batch: table {k!r} must be contiguous {dt} of shape {shp}")
    out = {k: torch.empty((b,) + shp[1:], device=dev, dtype=dt) for k, (dt, shp) in want.items()}
    bad = torch.zeros(1, device=dev, dtype=torch.int32)
    L.call("atmonr_gather_batch", *(L.ptr(tables[k]) for k in ("origin", "dir", "alt", "rad", "len", "idx", "irgb_idx")),
           L.ptr(index), b, r, *(L.ptr(out[k]) for k in ("origin", "dir", "alt", "rad", "len", "idx", "irgb_idx")),
           L.ptr(bad), L.stream())
    if gather_batch.check and int(bad.item()):
        raise IndexError("batch index out of range for the ray table")
    return out


gather_batch.check = False


# ------------------------------------------------------------------------------------------
# AtmoNeRF dense layers on tcgen05 (float32-accurate bf16x3 split), csrc/linear_tc.cu
# ------------------------------------------------------------------------------------------
# bf16 terms per float32 operand of the tensor-core dense layers (include/atmonr_b200.h): 2 = three partial
# products, error ~2^-16 of a product (the training default, checked against the float32 oracle at 1e-3 on the
# radiances), 3 = six partial products, float32-exact (cross-check; ATMONR_LINEAR_TERMS=3)
LINEAR_TERMS = int(os.environ.get("ATMONR_LINEAR_TERMS", "2"))


def linear_planes_bytes(n_out: int, k_in: int, terms: int | None = None) -> int:
    """Size of the split copy of a (n_out, k_in) matrix (include/atmonr_b200.h: atmonr_linear_prep)."""
    return -(-n_out // 256) * -(-k_in // 32) * (terms or LINEAR_TERMS) * 16384


def _rows(t: torch.Tensor):
    """(tensor, row stride in elements) of a 2-D float32 CUDA tensor whose rows are contiguous; a
    copy is made only when the layout does not allow that."""
    if not t.is_cuda:
        raise L.NativeLibraryError("libatmonr_b200 operates on CUDA tensors only (no CPU fallback)")
    if t.dtype != _f32 or t.dim() != 2 or t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.to(_f32).contiguous()
    return t, (t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0)))


def _segments(x: torch.Tensor, x2):
    """(x, ldx, x2 pointer, ldx2, k_split, k_in) for an input given as one tensor or as two column
    blocks [x | x2] that the kernels read in place (no torch.cat)."""
    x, ldx = _rows(x)
    if x2 is None:
        return x, ldx, None, None, 0, x.shape[1], x.shape[1]
    x2, ldx2 = _rows(x2)
    if x2.shape[0] != x.shape[0]:
        raise ValueError("the two input blocks must have the same number of rows")
    if x.shape[1] % 8:                                   # the kernels split on a multiple of 8 columns
        x, x2 = torch.cat([x, x2], dim=1), None
        return x, x.stride(0), None, None, 0, x.shape[1], x.shape[1]
    return x, ldx, x2, x2.data_ptr(), ldx2, x.shape[1], x.shape[1] + x2.shape[1]


def linear_forward(x: torch.Tensor, weight: torch.Tensor, bias, relu: bool, transpose: bool = False,
                   mask: torch.Tensor | None = None, x2: torch.Tensor | None = None, terms: int | None = None,
                   out_mask: torch.Tensor | None = None, out: torch.Tensor | None = None,
                   out_bits: torch.Tensor | None = None, want_bits: bool = False):
    """act(X' @ B.T + bias) with B = weight (n_out, k_in), or B = weight.T when `transpose` (then
    weight is (k_in, n_out)); X = x or [x | x2]; X' = X where mask > 0 and 0 elsewhere when a mask is
    given; the result is zeroed where out_mask (m, n_out) <= 0 when that is given. x, x2, mask and out_mask
    may be column slices of wider float32 tensors; `out`: a (m, >= n_out) float32 tensor (or column slice
    starting at a multiple of 4) that receives the result instead of a new tensor. want_bits (with relu,
    n_out % 128 == 0): also return the sign bits of the result, (m, n_out / 32) int32 in the kernel's own
    layout; out_bits: such an array, used like out_mask (32 bytes per 256-wide row instead of 1 KB)."""
    if not weight.is_cuda:
        raise L.NativeLibraryError("libatmonr_b200 operates on CUDA tensors only (no CPU fallback)")
    x, ldx, x2, x2p, ldx2, k_split, k_cols = _segments(x, x2)
    weight = _c(weight.detach(), _f32)
    n_out, k_in = (weight.shape[1], weight.shape[0]) if transpose else (weight.shape[0], weight.shape[1])
    if k_cols != k_in:
        raise ValueError(f"linear_forward: the input has {k_cols} columns, the matrix expects {k_in}")
    m = x.shape[0]
    mp, ldm = None, 0
    if mask is not None:
        mask, ldm = _rows(mask)
        if tuple(mask.shape) != (m, k_in):
            raise ValueError("linear_forward: mask and input must have the same shape")
        mp = mask.data_ptr()
    terms = terms or LINEAR_TERMS
    omp, ldom = None, 0
    if out_mask is not None:
        out_mask, ldom = _rows(out_mask)
        if out_mask.shape[0] != m or out_mask.shape[1] < n_out:
            raise ValueError("linear_forward: out_mask must be (rows, >= n_out)")
        omp = out_mask.data_ptr()
    planes = torch.empty(linear_planes_bytes(n_out, k_in, terms), device=x.device, dtype=torch.uint8)
    if out is None:
        y = torch.empty((m, n_out), device=x.device, dtype=_f32)
        yp, ldy = y.data_ptr(), n_out
    else:
        if out.dtype != _f32 or out.dim() != 2 or out.shape[0] != m or out.shape[1] < n_out or out.stride(1) != 1:
            raise ValueError("linear_forward: `out` must be a (rows, >= n_out) float32 tensor with contiguous rows")
        y, yp, ldy = out, out.data_ptr(), out.stride(0)
    bits = torch.empty((m, n_out // 32), device=x.device, dtype=torch.int32) if want_bits else None
    if out_bits is not None and tuple(out_bits.shape) != (m, n_out // 32):
        raise ValueError("linear_forward: out_bits must be (rows, n_out / 32)")
    L.call("atmonr_linear_prep", L.ptr(weight), n_out, k_in, int(transpose), terms, L.ptr(planes), L.stream())
    b = None if bias is None else _c(bias.detach(), _f32)
    L.call("atmonr_linear_fwd_tc", x.data_ptr(), ldx, x2p, ldx2, k_split, mp, ldm, L.ptr(planes), L.ptr(b), m, n_out,
           k_in, int(relu), terms, omp, ldom, L.ptr(out_bits), L.ptr(bits), yp, ldy, L.stream())
    return (y, bits) if want_bits else y


def linear_weight_grad(dy: torch.Tensor, x: torch.Tensor, mask: torch.Tensor | None = None,
                       x2: torch.Tensor | None = None, want_bias: bool = False, terms: int | None = None):
    """dW (n_out, k_in) = dy'.T @ X over all rows; X = x or [x | x2]; dy' = dy where mask > 0 (mask: the
    layer's output). With want_bias also db (n_out) = column sums of dy', from the same pass: (dW, db)."""
    dy, ldy = _rows(dy)
    x, ldx, x2, x2p, ldx2, k_split, k_in = _segments(x, x2)
    if dy.shape[0] != x.shape[0]:
        raise ValueError("linear_weight_grad: dy and x must have the same number of rows")
    mp, ldm = None, 0
    if mask is not None:
        mask, ldm = _rows(mask)
        if mask.shape != dy.shape:
            raise ValueError("linear_weight_grad: mask and dy must have the same shape")
        mp = mask.data_ptr()
    n_out = dy.shape[1]
    dw = torch.zeros((n_out, k_in), device=x.device, dtype=_f32)
    db = torch.zeros((n_out,), device=x.device, dtype=_f32) if want_bias else None
    L.call("atmonr_linear_dw_tc", dy.data_ptr(), ldy, mp, ldm, x.data_ptr(), ldx, x2p, ldx2, k_split, x.shape[0], n_out,
           k_in, terms or LINEAR_TERMS, L.ptr(dw), L.ptr(db), L.stream())
    return (dw, db) if want_bias else dw


class LinearTcFn(torch.autograd.Function):
    """torch.nn.functional.linear (+ optional ReLU) of x or of [x | x2] on the tensor cores: forward,
    input gradient and weight gradient are tcgen05 products (the ReLU derivative is applied to the
    incoming gradient while it is staged); the bias gradient comes out of the weight-gradient kernel."""

    @staticmethod
    def forward(ctx, x, x2, weight, bias, relu):
        y = linear_forward(x, weight, bias, relu, x2=x2)
        ctx.relu, ctx.has_bias, ctx.k1 = relu, bias is not None, x.shape[1]
        ctx.save_for_backward(x, x2, weight, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, x2, weight, y = ctx.saved_tensors
        dy = _c(dy, _f32)
        need_x = ctx.needs_input_grad[0] or (x2 is not None and ctx.needs_input_grad[1])
        dx = linear_forward(dy, weight, None, False, transpose=True, mask=y) if need_x else None
        dx1 = dx2 = None
        if dx is not None:
            dx1 = (dx if x2 is None else dx[:, : ctx.k1]) if ctx.needs_input_grad[0] else None
            dx2 = dx[:, ctx.k1:] if (x2 is not None and ctx.needs_input_grad[1]) else None
        want_db = ctx.has_bias and ctx.needs_input_grad[3]
        dw = db = None
        if ctx.needs_input_grad[2] or want_db:
            dw, db = linear_weight_grad(dy, x, mask=y, x2=x2, want_bias=True)
        return dx1, dx2, (dw if ctx.needs_input_grad[2] else None), (db if want_db else None), None


def linear_tc(x, weight, bias=None, relu: bool = False, x2=None):
    """relu?(cat([x, x2]) @ weight.T + bias) without materialising the concatenation."""
    return LinearTcFn.apply(x, x2, weight, bias, relu)


# ------------------------------------------------------------------------------------------
# hash grid
# ------------------------------------------------------------------------------------------
def hashgrid_indices(grid: L.GridT, x: torch.Tensor) -> torch.Tensor:
    x = _c(x, _f32)
    m = x.shape[0]
    idx = torch.empty((m, grid.n_levels, 1 << grid.n_dims), device=x.device, dtype=torch.int32)
    L.call("atmonr_hashgrid_indices", C.byref(grid), L.ptr(x), x.shape[1], m, L.ptr(idx), L.stream())
    return idx


class HashGridFn(torch.autograd.Function):
    """tcnn.Encoding(HashGrid).forward: x (M, >=D) float32, params flat float32."""

    @staticmethod
    def forward(ctx, x, params, table_f16, grid):
        x = _c(x, _f32)
        m = x.shape[0]
        out = torch.empty((m, 2 * grid.n_levels), device=x.device, dtype=_f32)
        L.call("atmonr_hashgrid_fwd", C.byref(grid), L.ptr(x), x.shape[1], L.ptr(table_f16), m, L.ptr(out), L.stream())
        ctx.save_for_backward(x)
        ctx.grid, ctx.n_params = grid, params.numel()
        return out

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        dtable = torch.zeros(ctx.n_params, device=x.device, dtype=_f32)
        L.call("atmonr_hashgrid_bwd", C.byref(ctx.grid), L.ptr(x), x.shape[1], L.ptr(_c(dout, _f32)), x.shape[0],
               L.ptr(dtable), L.stream())
        return None, dtable, None, None


# ------------------------------------------------------------------------------------------
# MLP
# ------------------------------------------------------------------------------------------
class MlpFn(torch.autograd.Function):
    """tcnn.Network.forward: x (M, n_in) float32 -> (M, n_out) float32."""

    @staticmethod
    def forward(ctx, x, params, w_f16, shape):
        x = _c(x, _f32)
        m = x.shape[0]
        out = torch.empty((m, shape.n_out), device=x.device, dtype=_f32)
        L.call("atmonr_mlp_fwd", C.byref(shape), L.ptr(w_f16), L.ptr(x), m, L.ptr(out), L.stream())
        ctx.save_for_backward(x, w_f16)
        ctx.shape, ctx.need_dx = shape, x.requires_grad
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w_f16 = ctx.saved_tensors
        shape = ctx.shape
        dw = torch.zeros(shape.n_params, device=x.device, dtype=_f32)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        L.call("atmonr_mlp_bwd", C.byref(shape), L.ptr(w_f16), L.ptr(x), L.ptr(_c(dout, _f32)), x.shape[0],
               L.ptr(dx), L.ptr(dw), L.stream())
        return dx, dw, None, None


# ------------------------------------------------------------------------------------------
# compositing
# ------------------------------------------------------------------------------------------
def composite_forward(z, color, sigma, color_surf, z_scale, relu, want_weights=True, want_alpha=True):
    b, n = z.shape
    k, v = color.shape[-1], sigma.shape[-1]
    dev = z.device
    cmap = torch.empty((b, k), device=dev, dtype=_f32)
    catmo = torch.empty((b, k), device=dev, dtype=_f32)
    csurf = torch.empty((b, k), device=dev, dtype=_f32)
    tsurf = torch.empty((b, v), device=dev, dtype=_f32)
    weights = torch.empty((b, n, v), device=dev, dtype=_f32) if want_weights else None
    alpha = torch.empty((b, n, v), device=dev, dtype=_f32) if want_alpha else None
    L.call("atmonr_composite_fwd", L.ptr(z), L.ptr(color), L.ptr(sigma), L.ptr(color_surf), float(z_scale), b, n, k, v,
           int(relu), L.ptr(cmap), L.ptr(catmo), L.ptr(csurf), L.ptr(tsurf), L.ptr(weights), L.ptr(alpha), L.stream())
    return cmap, catmo, csurf, tsurf, weights, alpha


def composite_backward(z, color, sigma, color_surf, catmo, tsurf, d_atmo, d_surf, z_scale, relu, want_dz=False,
                       grad_absmax=None, weights=None, d_weights=None):
    """atmonr_composite_bwd; with `d_weights` (dL/d of the per-sample weights the forward returned, NeRF
    coarse pass) atmonr_composite_bwd_weights. want_dz: also dL/dz (atmonr_composite_dz)."""
    b, n = z.shape
    k, v = color.shape[-1], sigma.shape[-1]
    dcolor = torch.empty_like(color)
    dsigma = torch.empty_like(sigma)
    dcs = torch.empty_like(color_surf) if color_surf is not None else None
    ddelta = torch.empty_like(z) if want_dz else None
    if d_weights is not None:
        L.call("atmonr_composite_bwd_weights", L.ptr(z), L.ptr(color), L.ptr(sigma), L.ptr(color_surf), L.ptr(catmo),
               L.ptr(tsurf), L.ptr(weights), L.ptr(d_atmo), L.ptr(d_surf), L.ptr(d_weights), float(z_scale), b, n, k, v,
               int(relu), L.ptr(dcolor), L.ptr(dsigma), L.ptr(dcs), L.ptr(ddelta), L.stream())
    else:
        L.call("atmonr_composite_bwd", L.ptr(z), L.ptr(color), L.ptr(sigma), L.ptr(color_surf), L.ptr(catmo),
               L.ptr(tsurf), L.ptr(d_atmo), L.ptr(d_surf), float(z_scale), b, n, k, v, int(relu), L.ptr(dcolor),
               L.ptr(dsigma), L.ptr(dcs), L.ptr(ddelta), L.ptr(grad_absmax), L.stream())
    if not want_dz:
        return dcolor, dsigma, dcs
    dz = torch.empty_like(z)
    L.call("atmonr_composite_dz", L.ptr(ddelta), b, n, float(z_scale), L.ptr(dz), L.stream())
    return dcolor, dsigma, dcs, dz


def composite_backward_compact(z, color, sigma, color_surf, catmo, tsurf, d_atmo, d_surf, z_scale, relu,
                               grad_absmax=None):
    """atmonr_composite_bwd_compact: gradients of the samples that can carry one, as a list.
    -> (active_idx (M,) int32 [first n valid], n_active (1,) int32 on the device, dcolor_c (M,K), dsigma_c (M,V),
    dcolor_surf)."""
    b, n = z.shape
    k, v = color.shape[-1], sigma.shape[-1]
    m = b * n
    active_idx = torch.empty(m, device=z.device, dtype=torch.int32)
    n_active = torch.zeros(1, device=z.device, dtype=torch.int32)
    dcolor_c = torch.empty((m, k), device=z.device, dtype=_f32)
    dsigma_c = torch.empty((m, v), device=z.device, dtype=_f32)
    dcs = torch.empty_like(color_surf) if color_surf is not None else None
    L.call("atmonr_composite_bwd_compact", L.ptr(z), L.ptr(color), L.ptr(sigma), L.ptr(color_surf), L.ptr(catmo),
           L.ptr(tsurf), L.ptr(d_atmo), L.ptr(d_surf), float(z_scale), b, n, k, v, int(relu), L.ptr(active_idx),
           L.ptr(n_active), L.ptr(dcolor_c), L.ptr(dsigma_c), L.ptr(dcs), L.ptr(grad_absmax), L.stream())
    return active_idx, n_active, dcolor_c, dsigma_c, dcs


class CompositeFn(torch.autograd.Function):
    """graphics_utils.py render / render_with_surface, differentiable w.r.t. colour, density,
    surface colour, the sample distances z (NeRF fine pass) and THROUGH the returned per-sample weights
    (NeRF coarse pass: they define the fine sampler's CDF, samplers.py:72-74)."""

    @staticmethod
    def forward(ctx, z, color, sigma, color_surf, z_scale, relu):
        z, color, sigma = _c(z, _f32), _c(color, _f32), _c(sigma, _f32)
        cs = _c(color_surf, _f32) if color_surf is not None else None
        cmap, catmo, csurf, tsurf, weights, alpha = composite_forward(z, color, sigma, cs, z_scale, relu)
        ctx.save_for_backward(z, color, sigma, cs, catmo, tsurf, weights)
        ctx.z_scale, ctx.relu = z_scale, relu
        ctx.mark_non_differentiable(alpha)
        return cmap, catmo, csurf, weights, alpha

    @staticmethod
    def backward(ctx, g_map, g_atmo, g_surf, g_w, _ga):
        z, color, sigma, cs, catmo, tsurf, weights = ctx.saved_tensors
        zero = torch.zeros_like(catmo)
        g_map = zero if g_map is None else g_map
        d_atmo = _c(g_map + (g_atmo if g_atmo is not None else 0), _f32)
        d_surf = _c(g_map + (g_surf if g_surf is not None else 0), _f32)
        out = composite_backward(z, color, sigma, cs, catmo, tsurf, d_atmo, d_surf, ctx.z_scale, ctx.relu,
                                 want_dz=ctx.needs_input_grad[0], weights=weights,
                                 d_weights=_c(g_w, _f32) if g_w is not None else None)
        dz = out[3] if ctx.needs_input_grad[0] else None
        return dz, out[0], out[1], out[2], None, None


# ------------------------------------------------------------------------------------------
# loss
# ------------------------------------------------------------------------------------------
class BandLossFn(torch.autograd.Function):
    """instant_ngp.py:249-263: select the ray's band, apply losses.py:<kind>, mean over rays.
    The gradient w.r.t. the colour map is produced by the same kernel."""

    @staticmethod
    def forward(ctx, color_map, band, rad, max_i, kind):
        cm = _c(color_map, _f32)
        b, k = cm.shape
        loss = torch.empty(1, device=cm.device, dtype=_f32)
        dcm = torch.empty_like(cm)
        partial = torch.empty(1024, device=cm.device, dtype=_f32)
        L.call("atmonr_band_loss", L.ptr(cm), L.ptr(_c(band, torch.int64)), L.ptr(_c(rad, _f32)), float(max_i), int(kind),
               b, k, 1.0, L.ptr(loss), L.ptr(dcm), L.ptr(partial), L.stream())
        ctx.save_for_backward(dcm)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (dcm,) = ctx.saved_tensors
        return dcm * g, None, None, None, None


def band_loss(color_map, band, rad, max_i, kind: str):
    return BandLossFn.apply(color_map, band, rad, max_i, LOSS_KINDS[kind])


# ------------------------------------------------------------------------------------------
# optimizer
# ------------------------------------------------------------------------------------------
def adamw_step(param, grad, exp_avg, exp_avg_sq, param_f16, lr, beta1, beta2, eps, weight_decay, step,
               grad_scale=1.0, zero_grad=False):
    L.call("atmonr_adamw_step", L.ptr(param), L.ptr(grad), L.ptr(exp_avg), L.ptr(exp_avg_sq), L.ptr(param_f16),
           param.numel(), float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), int(step),
           float(grad_scale), int(zero_grad), L.stream())


# ------------------------------------------------------------------------------------------
# NeRF helpers
# ------------------------------------------------------------------------------------------
def positional_encoding(pts: torch.Tensor, freqs, interleaved: bool) -> torch.Tensor:
    """encoders.py:4-28 -> float32 features. float64 points keep their precision through the phase
    (the extract path, nerf.py:209-213); everything else is encoded in float32."""
    f64 = pts.dtype == torch.float64
    p = _c(pts, torch.float64 if f64 else _f32)
    c = p.shape[-1]
    flat = p.reshape(-1, c)
    fl = list(freqs) if not isinstance(freqs, int) else [freqs] * c
    arr = (C.c_int32 * len(fl))(*fl)
    out = torch.empty((flat.shape[0], 2 * sum(fl)), device=p.device, dtype=_f32)
    L.call("atmonr_positional_encoding_f64" if f64 else "atmonr_positional_encoding", L.ptr(flat), flat.shape[0], c,
           arr, int(interleaved), L.ptr(out), L.stream())
    return out


def sample_pdf_z(weights, z_coarse, u):
    """samplers.py:72-101 -> (z_sorted, inds)."""
    w, zc, u = _c(weights, _f32), _c(z_coarse, _f32), _c(u, _f32)
    b, nc = zc.shape
    nf = u.shape[1]
    z = torch.empty((b, nc + nf), device=zc.device, dtype=_f32)
    inds = torch.empty((b, nf), device=zc.device, dtype=torch.int64)
    L.call("atmonr_sample_pdf", L.ptr(w), L.ptr(zc), L.ptr(u), b, nc, nf, L.ptr(z), L.ptr(inds), L.stream())
    return z, inds


class InverseCdfFn(torch.autograd.Function):
    """samplers.py:72-101: (weights (B,Nc), z_coarse (B,Nc), u (B,Nf)) -> (z_sorted (B,Nc+Nf), inds, cdf).
    Forward atmonr_sample_pdf_train, backward atmonr_sample_pdf_bwd (the reference's graph: only the bin
    width is detached, samplers.py:96, so gradients reach the coarse weights through the CDF)."""

    @staticmethod
    def forward(ctx, weights, z_coarse, u):
        w, zc, u = _c(weights, _f32), _c(z_coarse, _f32), _c(u, _f32)
        b, nc = zc.shape
        nf = u.shape[1]
        z = torch.empty((b, nc + nf), device=zc.device, dtype=_f32)
        inds = torch.empty((b, nf), device=zc.device, dtype=torch.int64)
        cdf = torch.empty((b, nc - 1), device=zc.device, dtype=_f32)
        src = torch.empty((b, nc + nf), device=zc.device, dtype=torch.int32)
        L.call("atmonr_sample_pdf_train", L.ptr(w), L.ptr(zc), L.ptr(u), b, nc, nf, L.ptr(z), L.ptr(inds), L.ptr(cdf),
               L.ptr(src), L.stream())
        ctx.save_for_backward(w, zc, u, cdf, inds, src)
        ctx.mark_non_differentiable(inds, cdf)
        return z, inds, cdf

    @staticmethod
    def backward(ctx, gz, _gi, _gc):
        w, zc, u, cdf, inds, src = ctx.saved_tensors
        b, nc = zc.shape
        nf = u.shape[1]
        dw = torch.empty_like(w) if ctx.needs_input_grad[0] else None
        dzc = torch.empty_like(zc) if ctx.needs_input_grad[1] else None
        L.call("atmonr_sample_pdf_bwd", L.ptr(_c(gz, _f32)), L.ptr(src), L.ptr(w), L.ptr(zc), L.ptr(u), L.ptr(cdf),
               L.ptr(inds), b, nc, nf, L.ptr(dw), L.ptr(dzc), L.stream())
        return dw, dzc, None


class NerfEncodeFn(torch.autograd.Function):
    """pipelines/nerf.py:104-135 for one pass: z (B,N) -> x (B*N, 2 sum(L_x) + 6 L_d) = [encoded preprocessed
    point | encoded ray direction] and the preprocessed points (B*N,3). One kernel forward
    (atmonr_nerf_encode); backward = dL/dz in one kernel (atmonr_nerf_encode_bwd: encoding derivative,
    float64 geodetic Jacobian, projection on the direction)."""

    @staticmethod
    def forward(ctx, z, origin, direction, frame, pos_freqs, dir_freqs):
        z, o, d = _c(z, _f32), _c(origin, _f32), _c(direction, _f32)
        b, n = z.shape
        fr = (C.c_int32 * 3)(*pos_freqs)
        width = 2 * sum(pos_freqs) + 6 * dir_freqs
        ld = (width + 3) // 4 * 4        # 16-byte aligned rows for the dense layers' vector loads
        x = torch.empty((b * n, ld), device=z.device, dtype=_f32)
        pts_n = torch.empty((b * n, 3), device=z.device, dtype=_f32)
        L.call("atmonr_nerf_encode", C.byref(frame), L.ptr(o), L.ptr(d), L.ptr(z), b, n, fr, int(dir_freqs), L.ptr(x), ld,
               L.ptr(pts_n), L.stream())
        ctx.frame, ctx.pos_freqs = frame, tuple(pos_freqs)
        ctx.save_for_backward(z, o, d, pts_n)
        ctx.mark_non_differentiable(pts_n)
        return (x if ld == width else x[:, :width]), pts_n

    @staticmethod
    def backward(ctx, gx, _gp):
        if not ctx.needs_input_grad[0]:
            return None, None, None, None, None, None
        z, o, d, pts_n = ctx.saved_tensors
        b, n = z.shape
        gx, ldg = _rows(gx)
        gz = torch.empty_like(z)
        fr = (C.c_int32 * 3)(*ctx.pos_freqs)
        L.call("atmonr_nerf_encode_bwd", C.byref(ctx.frame), L.ptr(o), L.ptr(d), L.ptr(z), L.ptr(pts_n), gx.data_ptr(), ldg,
               b, n, fr, L.ptr(gz), L.stream())
        return gz, None, None, None, None, None


def append_heights(pts: torch.Tensor, ray_origin_height: float, scale: float, offset) -> torch.Tensor:
    """samplers.py:168-195 on the device: (..., 3) float32 -> (..., 4)."""
    p = _c(pts, _f32)
    flat = p.reshape(-1, 3)
    out = torch.empty((flat.shape[0], 4), device=p.device, dtype=_f32)
    off = (C.c_double * 3)(*[float(v) for v in offset])
    L.call("atmonr_append_heights", L.ptr(flat), flat.shape[0], float(scale), off, float(ray_origin_height), L.ptr(out),
           L.stream())
    return out.view(*pts.shape[:-1], 4)
