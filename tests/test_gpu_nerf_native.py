"""GPU parity of the NeRF training path's kernels (csrc/nerf_points.cu, atmonr_composite_bwd_weights):
forward values and GRADIENTS against the oracle's torch graphs (which are pinned to the reference's
NeRFPipeline, tests/test_oracle_golden.py) and against the reference's golden vectors."""

import os

import numpy as np
import pytest
import torch

from helpers import take, tiny_scene
from oracle import geodesy, nerf as onerf, rendering, sampling

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))
T = lambda k: torch.from_numpy(G[k])


def _ops():
    from atmonr.native import lib as L, ops
    L.load()
    return L, ops


def _oracle_cdf(w):
    w = w[:, 1:-1]
    pdf = (w + 1e-8) / torch.sum(w + 1e-8, dim=1, keepdim=True)
    cdf = torch.cumsum(pdf, dim=1)
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=1)


@pytest.mark.parametrize("b,nc,nf", [(40, 64, 128), (7, 16, 24), (300, 8, 16)])
def test_sample_pdf_bin_indices_are_exact_and_gradients_match(b, nc, nf):
    """samplers.py:72-101. Bin indices: EXACTLY searchsorted(cdf, u, right=True) on the CDF the kernel
    built (north_star: bit-exact sample bin indices given identical uniforms and CDF); that CDF is the
    oracle's to a few float32 ulps (torch's CPU sum / cumsum orders are not reproducible across vector
    widths), and wherever u is further than that from a CDF edge the indices equal the oracle's.
    Gradients w.r.t. the coarse weights (through the CDF) and the coarse distances: the oracle's autograd."""
    L, ops = _ops()
    g = torch.Generator().manual_seed(51 + nc)
    w = torch.rand(b, nc, generator=g)
    w[0, 3:6] = 0.0                       # empty bins: the `den < 1e-8 -> 1` branch of samplers.py:92
    if b > 1:
        w[1] = 0.0                        # a ray without any weight: uniform CDF from the 1e-8 floor
    zc = torch.sort(torch.rand(b, nc, generator=g), dim=1)[0]
    u = torch.rand(b, nf, generator=g)
    wo, zo = w.clone().requires_grad_(), zc.clone().requires_grad_()
    z_o, inds_o = sampling.inverse_cdf_z(wo[..., None], zo, u)
    wd, zd = w.cuda().requires_grad_(), zc.cuda().requires_grad_()
    z, inds, cdf = ops.InverseCdfFn.apply(wd, zd, u.cuda())
    assert bool((torch.diff(z, dim=1) >= 0).all())
    assert torch.allclose(z.detach().cpu(), z_o.detach(), atol=1e-5)
    assert torch.equal(inds, torch.searchsorted(cdf, u.cuda().contiguous(), right=True))
    cdf_o = _oracle_cdf(w)
    assert float((cdf.cpu() - cdf_o).abs().max()) <= 4e-7      # a few ulps: the total's last bit moves every entry
    gap = (u[:, :, None] - cdf_o[:, None, :]).abs().min(dim=2)[0]
    clear = gap > 1e-6
    assert float(clear.float().mean()) > 0.99 and torch.equal(inds.cpu()[clear], inds_o[clear])
    gz = torch.randn(b, nc + nf, generator=g)
    (z_o * gz).sum().backward()
    (z * gz.cuda()).sum().backward()
    for got, want in ((wd.grad.cpu(), wo.grad), (zd.grad.cpu(), zo.grad)):
        assert float(want.abs().max()) > 0
        assert float((got - want).abs().max()) <= 2e-4 * float(want.abs().max())
    assert float(wd.grad[:, 0].abs().max()) == 0 and float(wd.grad[:, -1].abs().max()) == 0


def test_sample_pdf_matches_the_reference_vectors():
    """The reference's own sample_pdf on torch.manual_seed(77) draws (tests/golden/make_golden.py)."""
    L, ops = _ops()
    from atmonr import samplers
    z, inds, _ = ops.InverseCdfFn.apply(T("pdf_w")[..., 0].cuda(), T("pdf_zc").cuda(), T("pdf_u").cuda())
    assert torch.allclose(z.cpu(), T("pdf_z"), rtol=1e-6, atol=1e-6)
    o, d = T("rays_origin_norm")[::3].contiguous(), T("rays_dir")[::3].contiguous()
    batch = {"origin": o[:7].cuda(), "dir": d[:7].cuda()}
    torch.manual_seed(77)
    torch.cuda.manual_seed(77)
    pts, z2 = samplers.sample_pdf(batch, T("pdf_w").cuda(), T("pdf_zc").cuda(), n_samples=24)
    assert pts.shape == (7, 40, 3) and z2.shape == (7, 40) and bool((torch.diff(z2, dim=1) >= 0).all())


@pytest.mark.parametrize("v", [1, 4])
def test_compositing_is_differentiable_through_its_weights(v):
    """graphics_utils.py:28-48: a loss on the colour map AND on the returned per-sample weights (the NeRF
    coarse pass feeds them to sample_pdf) and on z: gradients of colour, density and distances against
    autograd through the oracle's cumprod graph."""
    L, ops = _ops()
    from atmonr.graphics_utils import render
    g = torch.Generator().manual_seed(70 + v)
    b, n, k = 33, 77, 4
    z = torch.sort(torch.rand(b, n, generator=g), dim=1)[0] * 30
    col = torch.rand(b, n, k, generator=g)
    sg = torch.rand(b, n, v, generator=g) * 0.2
    sg[:, 5:9] = 0.0
    gw = torch.randn(b, n, v, generator=g)
    gc = torch.randn(b, k, generator=g)

    def run(dev, fn):
        zz, cc, ss = (t.to(dev).clone().requires_grad_() for t in (z, col, sg))
        cmap, _, wts = fn(zz, cc, ss)
        ((cmap * gc.to(dev)).sum() + (wts * gw.to(dev)).sum()).backward()
        return cmap.detach().cpu(), wts.detach().cpu(), zz.grad.cpu(), cc.grad.cpu(), ss.grad.cpu()

    want = run("cpu", rendering.composite)
    got = run("cuda", render)
    for a, bb in zip(got, want):
        assert float((a - bb).abs().max()) <= 2e-5 * max(1.0, float(bb.abs().max()))


@pytest.mark.parametrize("with_frame", [True, False])
def test_nerf_point_encoder_forward_and_dz(with_frame):
    """atmonr_nerf_encode / _bwd on the device against the oracle (see the host-build twin in
    tests/test_abi_and_host.py for the tolerances)."""
    L, ops = _ops()
    scene = tiny_scene()
    b = take(scene.batch, slice(0, 150))
    nb, n, lx, ld = 150, 24, [14, 14, 10], 4
    g = torch.Generator().manual_seed(9)
    z = (torch.rand(nb, n, generator=g) * b["len"][:, None]).contiguous()
    z[0, 0] = b["len"][0] * 1.5
    if with_frame:
        fr = scene.frame
        frame = L.make_frame(fr.scale, fr.offset, fr.lat_min, fr.lat_range, fr.lon_min, fr.lon_range, fr.origin_height, fr.shift_lon)
    else:
        frame = L.disabled_frame()
    zd = z.cuda().requires_grad_()
    x, pn = ops.NerfEncodeFn.apply(zd, b["origin"].cuda(), b["dir"].cuda(), frame, tuple(lx), ld)
    zt = z.clone().requires_grad_()
    pts = b["origin"][:, None] + b["dir"][:, None] * zt[..., None]
    ptn = geodesy.preprocess_horizontal(pts, scene.frame) if with_frame else pts
    want = torch.cat([onerf.pe_per_axis(ptn, lx).view(nb * n, -1),
                      onerf.pe_interleaved(b["dir"][:, None].repeat(1, n, 1), ld).view(nb * n, -1)], dim=1)
    assert x.shape == want.shape == (nb * n, 100)
    assert float((pn.cpu() - ptn.detach().view(-1, 3)).abs().max()) < 1.5e-7
    assert float((x.detach().cpu() - want.detach()).abs().max()) < 6e-3     # 2^13 pi times a last-bit difference
    assert float((x.detach().cpu()[:, :4] - want.detach()[:, :4]).abs().max()) < 2e-6
    gx = torch.randn(nb * n, 100, generator=g)
    (want * gx).sum().backward()
    (x * gx.cuda()).sum().backward()
    ref = zt.grad
    assert float((zd.grad.cpu() - ref).abs().max()) <= 2e-3 * float(ref.abs().max())
    # gradient arriving through COLUMN SLICES of x (the way fc1 / fc6 consume it)
    zd2 = z.cuda().requires_grad_()
    x2 = ops.NerfEncodeFn.apply(zd2, b["origin"].cuda(), b["dir"].cuda(), frame, tuple(lx), ld)[0]
    (x2[:, :76] * gx.cuda()[:, :76]).sum().backward()
    assert torch.allclose(zd2.grad, zd.grad, rtol=1e-5, atol=1e-6 * float(ref.abs().max()))


def test_append_heights_kernel_matches_oracle():
    """samplers.py:168-195: the reference's own output (golden) and the oracle on random points."""
    L, ops = _ops()
    from atmonr import samplers
    gx = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors_extra.npz"))
    ref = torch.from_numpy(gx["ah_out"])
    got = samplers.append_heights(torch.from_numpy(gx["ah_pts"]).cuda(), 20000.0, float(gx["ah_scale"]),
                                  torch.from_numpy(gx["ah_offset"]).cuda())
    assert got.shape == ref.shape and float((got.cpu() - ref).abs().max()) <= 2e-7 * float(ref[..., 3].abs().max())
    scene = tiny_scene()
    g = torch.Generator().manual_seed(4)
    pts = torch.rand(50, 9, 3, generator=g) * 2 - 1
    off = torch.tensor(scene.frame.offset, dtype=torch.float64)
    want = sampling.append_heights(pts, 20000.0, scene.frame.scale, off)
    got = samplers.append_heights(pts.cuda(), 20000.0, scene.frame.scale, off.cuda())
    assert got.shape == (50, 9, 4) and torch.equal(got[..., :3].cpu(), pts)
    assert float((got[..., 3].cpu() - want[..., 3]).abs().max()) <= 2e-7 * float(want[..., 3].abs().max())
