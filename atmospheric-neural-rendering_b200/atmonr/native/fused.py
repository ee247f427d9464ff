"""Fused Instant-NGP render step (forward and backward) on libatmonr_b200.

Forward (instant_ngp.py:129-192 of the reference):
    rays -> [sample + geodetic preprocess] -> [hash grid + pos_mlp + SH + dir_mlp per sample]
         -> [surface branch per ray] -> [compositing]  ->  colour maps (B, 4)
Backward: compositing backward -> field backward (recomputes the per-sample activations,
accumulates MLP weight gradients per CTA, scatters table gradients with vector REDs) ->
surface backward.

Per-sample tensors kept between forward and backward: x01 (12 B), z (4 B), raw sigma (4 B),
raw colour (16 B). The (B, N, .) outputs the reference returns eagerly are materialised only
when somebody reads them (LazyResults).
"""

from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import torch

from atmonr.native import lib as L
from atmonr.native import ops

_f32 = torch.float32

# "tc": dense layers on tcgen05 tensor cores (default); "simt": the thread-per-sample FMA kernels
# (same arithmetic contract; kept as the cross-check of the tensor-core path).
FIELD_IMPL = os.environ.get("ATMONR_FIELD_IMPL", "tc")
# keep the encoded features of the forward pass for the backward pass when they fit in this many bytes
# ATMONR_COMPACT_BWD=1: the field backward visits only the samples whose incoming gradient can be
# non-zero (raw density > 0; exact, see atmonr_composite_bwd_compact). Pays off on scenes that are
# mostly empty; on the benchmark's randomly initialised field 82 % of the samples are listed and the
# list costs about what it saves (device-resident step -1.5 %, end-to-end step +3 %), so it is opt-in.
COMPACT_BWD = os.environ.get("ATMONR_COMPACT_BWD", "0") != "0"
ENC_CACHE_BYTES = int(float(os.environ.get("ATMONR_ENC_CACHE_GB", "40")) * (1 << 30))


@dataclass
class NGPState:
    """Static description of one InstantNGPPipeline for the fused kernels."""

    frame: L.FrameT
    grid3: L.GridT
    grid2: L.GridT
    pos_mlp: L.MlpT
    dir_mlp: L.MlpT
    surf_mlp: L.MlpT
    n_samples: int
    alt_compress: float
    z_scale: float            # scale / 1000 (normalised distance -> km)
    n_density: int = 1        # 4 with `multi_band_extinction` (instant_ngp.py:46-50)
    height: tuple | None = None   # (scale, offset, ray_origin_height) with `include_height`: 4-D grid
    seed: int = 0
    step: int = 0             # advances once per forward: new stratified draws every step
    ray_index_base: int = 0   # global index of the first ray of this rank's shard
    bins: torch.Tensor | None = None
    last: dict = field(default_factory=dict)
    pending: list = field(default_factory=list)   # sample points of announced batches (schedule_prefetch)
    side_stream: torch.cuda.Stream | None = None
    grad_bufs: list | None = None                 # persistent gradient buffers (see _gradient_buffers)


# Persistent gradient buffers: data_ptr -> [clean]. The backward accumulates into buffers that live as long
# as the pipeline; FusedAdamW recognises them by address and asks the AdamW kernel to zero each gradient in
# the pass that consumes it (atmonr_adamw_step, zero_grad = 1), so no 191 MB memset runs per step. A buffer
# is only handed out while it is known to be zero ("clean"); otherwise (two backward passes without an
# optimizer step in between, a foreign optimizer) fresh zeroed tensors are used, exactly as before.
PERSISTENT_GRADS: dict[int, list] = {}


def _gradient_buffers(st: NGPState, sizes, dev):
    if st.grad_bufs is None or any(b.numel() != n or b.device != dev for b, n in zip(st.grad_bufs, sizes)):
        for b in st.grad_bufs or []:
            PERSISTENT_GRADS.pop(b.data_ptr(), None)
        st.grad_bufs = [torch.zeros(n, device=dev, dtype=_f32) for n in sizes]
        for b in st.grad_bufs:
            PERSISTENT_GRADS[b.data_ptr()] = [True]
    out = []
    for b, n in zip(st.grad_bufs, sizes):
        flag = PERSISTENT_GRADS[b.data_ptr()]
        if flag[0]:
            flag[0] = False
            # a fresh alias: autograd adopts a gradient it holds the only reference to without copying it
            out.append(b.view(-1))
        else:
            out.append(torch.zeros(n, device=dev, dtype=_f32))
    return out


def _sample_seed(st: NGPState) -> int:
    """New stratified draws for every sampler invocation (the kernel's generator is counter-based)."""
    st.step += 1
    return (st.seed * 0x9E3779B97F4A7C15 + st.step) & 0xFFFFFFFFFFFFFFFF


def _rays_f32(origin, direction, length):
    return origin.contiguous().float(), direction.contiguous().float(), length.contiguous().float()


def schedule_prefetch(st: NGPState, origin, direction, length) -> None:
    """Announce the NEXT batch (call it before the current step's forward is enqueued).

    The sampler depends on the rays only, not on the parameters, and is bound by the FP64 pipe,
    while the field backward leaves most issue slots idle. So the next batch's
    sample points are computed on a side stream underneath the current step's backward:
    this call allocates the outputs and records "rays ready" on the current stream; the kernel
    itself is launched by `launch_prefetch` right after the field backward has been enqueued.
    """
    o, d, ln = _rays_f32(origin, direction, length)
    b, n = o.shape[0], st.n_samples
    if st.bins is None or st.bins.device != o.device or st.bins.numel() != n:
        st.bins = ops.linspace_bins(n, o.device)
    x01 = torch.empty((b * n, 3 if st.height is None else 4), device=o.device, dtype=_f32)
    z = torch.empty((b, n), device=o.device, dtype=_f32)
    ready = torch.cuda.Event()
    ready.record()
    st.pending.append({"key": (origin.data_ptr(), tuple(origin.shape)), "src": origin, "rays": (o, d, ln),
                       "x01": x01, "z": z, "seed": _sample_seed(st), "ready": ready, "done": None,
                       "ray_index_base": st.ray_index_base})
    del st.pending[:-2]  # the batch about to run and the one after it; anything older was never used


def launch_prefetch(st: NGPState) -> None:
    """Launch the announced batches' samplers on the side stream. The host runs ahead of the device,
    so the kernel starts as soon as its rays are ready and fills whatever the kernels of the current
    step leave idle (the step is 1.1 % faster than with in-line sampling, measured)."""
    for p in st.pending:
        if p["done"] is None:
            _launch_one(st, p, True)


def _launch_one(st: NGPState, p: dict, on_side_stream: bool) -> None:
    o, d, ln = p["rays"]
    cur = torch.cuda.current_stream()
    if on_side_stream:
        if st.side_stream is None or st.side_stream.device != o.device:
            st.side_stream = torch.cuda.Stream(device=o.device)
        run_on = st.side_stream
        run_on.wait_event(p["ready"])
    else:
        run_on = cur
    with torch.cuda.stream(run_on):
        ops.ngp_sample_points(st.frame, o, d, ln, st.n_samples, st.alt_compress, random=True, seed=p["seed"],
                              ray_index_base=p["ray_index_base"], bins=st.bins, out=(p["x01"], p["z"]), height=st.height)
        p["done"] = torch.cuda.Event()
        p["done"].record()


def take_prefetched(st: NGPState, origin):
    """-> (x01, z) of a batch announced with schedule_prefetch, or None."""
    key = (origin.data_ptr(), tuple(origin.shape))
    for k, p in enumerate(st.pending):
        if p["key"] == key:
            break
    else:
        return None
    if p["done"] is None:            # no backward ran in between (first step, evaluation): sample in line
        _launch_one(st, p, False)
    else:
        torch.cuda.current_stream().wait_event(p["done"])
    del st.pending[:k + 1]
    return p["x01"], p["z"]


def field_impl(st: NGPState) -> str:
    """The tcgen05 kernels cover the shipped shape (3-D grid, one density); `include_height` (4-D grid) and
    `multi_band_extinction` (four densities) run the same launch chain on the thread-per-sample kernels,
    which are templated on both (csrc/atmonr_b200.cu: k_field_fwd / k_field_bwd <D, V>)."""
    return "simt" if (FIELD_IMPL == "simt" or st.n_density != 1 or st.height is not None) else "tc"


def field_forward(st: NGPState, table16, pos_w16, dir_w16, x01, dirs, b, n, want_enc=False):
    """-> (sigma_raw (M, V), color_raw (M,4), enc (M,32) fp16 | None)."""
    sigma_raw = torch.empty((b * n, st.n_density), device=x01.device, dtype=_f32)
    color_raw = torch.empty((b * n, 4), device=x01.device, dtype=_f32)
    if field_impl(st) == "simt":
        L.call("atmonr_ngp_field_fwd", C.byref(st.grid3), L.ptr(table16), C.byref(st.pos_mlp), L.ptr(pos_w16),
               C.byref(st.dir_mlp), L.ptr(dir_w16), L.ptr(x01), L.ptr(dirs), b, n, L.ptr(sigma_raw), L.ptr(color_raw),
               L.stream())
        return sigma_raw, color_raw, None
    enc = None
    if want_enc and b * n * 64 <= ENC_CACHE_BYTES:
        enc = torch.empty((b * n, 32), device=x01.device, dtype=torch.float16)
    L.call("atmonr_ngp_field_fwd_tc", C.byref(st.grid3), L.ptr(table16), C.byref(st.pos_mlp), L.ptr(pos_w16),
           C.byref(st.dir_mlp), L.ptr(dir_w16), L.ptr(x01), L.ptr(dirs), b, n, L.ptr(sigma_raw), L.ptr(color_raw),
           L.ptr(enc), L.stream())
    return sigma_raw, color_raw, enc


def surface_forward(st: NGPState, table16, w16, origin, direction, length):
    b = origin.shape[0]
    out = torch.empty((b, 4), device=origin.device, dtype=_f32)
    L.call("atmonr_ngp_surface_fwd", C.byref(st.grid2), L.ptr(table16), C.byref(st.surf_mlp), L.ptr(w16),
           L.ptr(origin), L.ptr(direction), L.ptr(length), b, L.ptr(out), L.stream())
    return out


class NGPRenderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos_table, pos_w, dir_w, surf_table, surf_w, st: NGPState, shadows, origin, direction, length, u):
        t16, pw16, dw16, s16, sw16 = shadows
        pre = take_prefetched(st, origin) if u is None else None
        origin, direction, length = _rays_f32(origin, direction, length)
        b, n = origin.shape[0], st.n_samples
        if pre is not None:
            x01, z = pre
        else:
            if st.bins is None or st.bins.device != origin.device or st.bins.numel() != n:
                st.bins = ops.linspace_bins(n, origin.device)
            x01, z = ops.ngp_sample_points(st.frame, origin, direction, length, n, st.alt_compress, u=u, random=True,
                                           seed=_sample_seed(st), ray_index_base=st.ray_index_base, bins=st.bins,
                                           height=st.height)
        needs_grad = any(ctx.needs_input_grad[:3])  # (grad mode is off inside Function.forward)
        sigma_raw, color_raw, enc = field_forward(st, t16, pw16, dw16, x01, direction, b, n, want_enc=needs_grad)
        cs_raw = surface_forward(st, s16, sw16, origin, direction, length)
        cmap, catmo, csurf, tsurf, _, _ = ops.composite_forward(
            z, color_raw.view(b, n, 4), sigma_raw.view(b, n, st.n_density), cs_raw, st.z_scale, relu=True,
            want_weights=False, want_alpha=False)
        ctx.st = st
        ctx.enc = enc
        ctx.save_for_backward(t16, pw16, dw16, s16, sw16, origin, direction, length, x01, z, sigma_raw, color_raw,
                              cs_raw, catmo, tsurf)
        ctx.sizes = (pos_table.numel(), pos_w.numel(), dir_w.numel(), surf_table.numel(), surf_w.numel())
        st.last = {"z": z, "sigma_raw": sigma_raw.view(b, n, st.n_density), "color_raw": color_raw.view(b, n, 4),
                   "color_surf_raw": cs_raw, "x01": x01}
        return cmap, catmo, csurf

    @staticmethod
    def backward(ctx, g_map, g_atmo, g_surf):
        st = ctx.st
        (t16, pw16, dw16, s16, sw16, origin, direction, length, x01, z, sigma_raw, color_raw, cs_raw, catmo,
         tsurf) = ctx.saved_tensors
        b, n = z.shape
        dev = z.device
        g_map = torch.zeros_like(catmo) if g_map is None else g_map
        d_atmo = (g_map + g_atmo if g_atmo is not None else g_map).contiguous().float()
        d_surf = (g_map + g_surf if g_surf is not None else g_map).contiguous().float()
        simt = field_impl(st) == "simt"
        v = st.n_density
        absmax = torch.zeros(1, device=dev, dtype=_f32) if not simt else None
        compact = COMPACT_BWD and not simt and ctx.enc is not None and os.environ.get("ATMONR_BWD_NARROW") is None
        if compact:
            act_idx, n_act, dcolor, dsigma, dcs = ops.composite_backward_compact(
                z, color_raw.view(b, n, 4), sigma_raw.view(b, n, v), cs_raw, catmo, tsurf, d_atmo, d_surf,
                st.z_scale, relu=True, grad_absmax=absmax)
            st.last["n_active"] = n_act
        else:
            dcolor, dsigma, dcs = ops.composite_backward(
                z, color_raw.view(b, n, 4), sigma_raw.view(b, n, v), cs_raw, catmo, tsurf, d_atmo, d_surf,
                st.z_scale, relu=True, grad_absmax=absmax)
        d_table, d_pw, d_dw, d_s, d_sw = _gradient_buffers(st, ctx.sizes, dev)
        if simt:
            L.call("atmonr_ngp_field_bwd", C.byref(st.grid3), L.ptr(t16), C.byref(st.pos_mlp), L.ptr(pw16),
                   C.byref(st.dir_mlp), L.ptr(dw16), L.ptr(x01), L.ptr(direction), L.ptr(dsigma), L.ptr(dcolor), b, n,
                   L.ptr(d_table), L.ptr(d_pw), L.ptr(d_dw), L.stream())
            launch_prefetch(st)
        elif compact:
            L.call("atmonr_ngp_field_bwd_tc_compact", C.byref(st.grid3), C.byref(st.pos_mlp), L.ptr(pw16),
                   C.byref(st.dir_mlp), L.ptr(dw16), L.ptr(x01), L.ptr(direction), L.ptr(ctx.enc), L.ptr(act_idx),
                   L.ptr(n_act), L.ptr(dsigma), L.ptr(dcolor), L.ptr(absmax), b, n, L.ptr(d_table), L.ptr(d_pw),
                   L.ptr(d_dw), L.stream())
            launch_prefetch(st)
            ctx.enc = None
        else:
            L.call("atmonr_ngp_field_bwd_tc", C.byref(st.grid3), L.ptr(t16), C.byref(st.pos_mlp), L.ptr(pw16),
                   C.byref(st.dir_mlp), L.ptr(dw16), L.ptr(x01), L.ptr(direction), L.ptr(ctx.enc), L.ptr(dsigma),
                   L.ptr(dcolor), L.ptr(absmax), b, n, L.ptr(d_table), L.ptr(d_pw), L.ptr(d_dw), L.stream())
            # announced batches: their samplers go to the side stream now. (Releasing them together
            # with the backward from a common event, the backward on a high-priority stream, was
            # measured too: the sampler's CTAs do not fit next to two backward CTAs -- the register
            # file of each SM sub-partition is full -- so it only ran after the backward. Running the
            # sampler on a tenth warp INSIDE the backward kernel fits the register file but made the
            # step 2-4 ms slower: its ~1600 FP64 instructions share the instruction cache with the
            # backward's 3900.)
            launch_prefetch(st)
            ctx.enc = None
        L.call("atmonr_ngp_surface_bwd", C.byref(st.grid2), L.ptr(s16), C.byref(st.surf_mlp), L.ptr(sw16),
               L.ptr(origin), L.ptr(direction), L.ptr(length), L.ptr(dcs), b, L.ptr(d_s), L.ptr(d_sw), L.stream())
        return d_table, d_pw, d_dw, d_s, d_sw, None, None, None, None, None, None


class LazyResults(dict):
    """Result dict of InstantNGPPipeline.forward. The (B, N, .) entries of the reference
    (instant_ngp.py:194-203) are computed on first access from the saved raw field outputs;
    they are detached (the trainer never differentiates through them)."""

    LAZY = ("color_fine", "sigma_fine", "weights_fine", "z_vals_fine", "color_surf")
    LAZY_HEIGHT = ("norm_heights_fine",)        # instant_ngp.py:204-205, `include_height` only

    def __init__(self, eager: dict, st: NGPState):
        super().__init__(eager)
        self._st = st
        self._buf = dict(st.last)
        if st.height is not None:
            self.LAZY = self.LAZY + self.LAZY_HEIGHT

    def _materialise(self, key):
        b = self._buf
        if key == "color_fine":
            return torch.relu(b["color_raw"])[:, :-1]
        if key == "sigma_fine":
            return torch.relu(b["sigma_raw"])[:, :-1]
        if key == "z_vals_fine":
            return b["z"]
        if key == "norm_heights_fine":
            return b["x01"][:, 3].view(b["z"].shape)
        if key == "color_surf":
            return torch.relu(b["color_surf_raw"])
        if key == "weights_fine":
            return ops.composite_forward(b["z"], b["color_raw"], b["sigma_raw"], b["color_surf_raw"],
                                         self._st.z_scale, relu=True, want_weights=True, want_alpha=False)[4]
        raise KeyError(key)

    def __getitem__(self, key):
        if not dict.__contains__(self, key) and key in self.LAZY:
            dict.__setitem__(self, key, self._materialise(key))
        return dict.__getitem__(self, key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self.LAZY

    def keys(self):
        return list(dict.keys(self)) + [k for k in self.LAZY if not dict.__contains__(self, k)]


def extract_sigma(st: NGPState, table16, pos_w16, pts: torch.Tensor) -> torch.Tensor:
    """instant_ngp.py:208-247: (n,3) float64 normalised points -> (n,1) float32 density."""
    p = pts.contiguous().double()
    n = p.shape[0]
    out = torch.empty(n, device=p.device, dtype=_f32)
    L.call("atmonr_extract_sigma" if field_impl(st) == "simt" else "atmonr_extract_sigma_tc", C.byref(st.frame), C.byref(st.grid3), L.ptr(table16), C.byref(st.pos_mlp),
           L.ptr(pos_w16), L.ptr(p), n, float(st.alt_compress), L.ptr(out), L.stream())
    return out.view(n, 1)
