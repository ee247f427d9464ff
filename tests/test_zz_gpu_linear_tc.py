"""GPU parity of the tcgen05 dense layer of the AtmoNeRF MLP (csrc/linear_tc.cu) against float64
matrix products, torch autograd and the default (library GEMM) NeRF pipeline.

Tolerances: every product is required to be as close to the float64 result as max(4e-6 of the output
scale [2e-5 for the row-reduction of the weight gradient], twice the error of the library float32
product of the same operands): "float32-grade", not a fixed ulp count, because the rounding of a K-
(or M-) term float32 sum grows with the number of terms (see _close). That is the three-term flavour
(terms = 3, six partial products). The two-term flavour the pipelines train with (terms = 2: three partial
products, operands carried to 16 significand bits) is held to 5e-5 of the output scale per product (2e-4
for the weight gradient), and the whole NeRF step to 3e-4 on the colour maps, inside north_star's 1e-3."""

import os

import pytest
import torch

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module", autouse=True)
def built():
    import __graft_entry__ as ge
    ge.build()
    assert torch.cuda.is_available()


SHAPES = [  # (M, k_in, n_out, relu, bias): the layer shapes of configs/nerf.json + ragged edges
    (300, 76, 256, True, True),      # fc1 (K tail: 76 = 2 chunks + 12)
    (1000, 256, 256, True, True),    # fc2-5, fc7-8
    (129, 332, 256, True, True),     # fc6 (skip connection)
    (257, 256, 257, False, True),    # fc9 coarse: 256 + 1 outputs -> second column tile of 16
    (64, 256, 260, False, True),     # fc9 fine
    (500, 280, 128, True, True),     # fc10
    (333, 128, 4, False, True),      # fc11
    (1, 32, 16, False, False), (128, 7, 3, True, False), (4096, 256, 256, False, False),
    (40000, 332, 260, True, True),   # more 32-row chunks than SMs: several chunks per CTA in the weight gradient
]


@pytest.fixture(params=[3, 2], ids=["terms3", "terms2"])
def terms(request, monkeypatch):
    from atmonr.native import ops
    monkeypatch.setattr(ops, "LINEAR_TERMS", request.param)
    return request.param


def _ref(x, w, b, relu):
    y = x.double() @ w.double().t() + (0 if b is None else b.double())
    return torch.relu(y) if relu else y


def _close(got, want64, lib32, what, terms=3):
    """float32-grade: within max(tol, 2 x the library float32 product's error) of the float64 result,
    relative to the output scale. tol = 4e-6 for chains of up to ~100 tensor-core accumulations (K <= 332:
    21 K-steps x 6 products); the TMEM accumulator TRUNCATES each float32 addition (measured: the error
    grows linearly with the number of accumulations, ~6e-8 each), so the weight gradient, which sums
    M / (16 x CTAs) K-steps per CTA, gets 2e-5."""
    scale = float(want64.abs().max()) + 1e-30
    err = float((got.double() - want64).abs().max()) / scale
    lib = float((lib32.double() - want64).abs().max()) / scale
    tol = 2e-5 if what in ("weight gradient", "bias gradient") else 4e-6
    if terms == 2 and what != "bias gradient":   # (the bias gradient is a float32 column sum, no split involved)
        tol = 2e-4 if what == "weight gradient" else 5e-5
    assert err <= max(tol, 2 * lib), (what, err, lib)


@pytest.mark.parametrize("m,k,n,relu,bias", SHAPES)
def test_linear_forward_matches_float64(m, k, n, relu, bias, terms):
    from atmonr.native import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(m + k + n)
    x = torch.randn(m, k, generator=g).cuda()
    w = (torch.randn(n, k, generator=g) / k ** 0.5).cuda()
    b = torch.randn(n, generator=g).cuda() if bias else None
    want = _ref(x, w, b, relu)
    got = ops.linear_forward(x, w, b, relu)
    lib = x @ w.t() + (0 if b is None else b)
    _close(got, want, torch.relu(lib) if relu else lib, "forward", terms)
    # input-gradient form: dY (m, n) * W (n, k) through the planes of W^T
    dy = torch.randn(m, n, generator=g).cuda()
    dx = ops.linear_forward(dy, w, None, False, transpose=True)
    _close(dx, dy.double() @ w.double(), dy @ w, "input gradient", terms)
    # the same with the ReLU derivative of the layer's output applied while dY is staged
    keep = (got > 0) if relu else torch.ones_like(got, dtype=torch.bool)
    dx = ops.linear_forward(dy, w, None, False, transpose=True, mask=got if relu else None)
    _close(dx, (dy * keep).double() @ w.double(), (dy * keep) @ w, "masked input gradient", terms)
    # weight gradient: a reduction over all rows (split over the CTAs, float32 REDs at the end)
    dw, db = ops.linear_weight_grad(dy, x, mask=got if relu else None, want_bias=True)
    assert dw.shape == (n, k) and db.shape == (n,)
    _close(dw, (dy * keep).double().t() @ x.double(), (dy * keep).t() @ x, "weight gradient", terms)
    _close(db, (dy * keep).double().sum(0), (dy * keep).sum(0), "bias gradient", terms)


def test_linear_on_a_column_slice_and_autograd(terms):
    from atmonr.native import ops
    ftol, gtol = (4e-6, 1e-5) if terms == 3 else (5e-5, 1e-4)
    g = torch.Generator().manual_seed(0)
    wide = torch.randn(700, 100, generator=g).cuda()
    x = wide[:, :76]                                   # models/nerf.py: x_pos = x[:, :pos_channels]
    w = (torch.randn(256, 76, generator=g) / 9).cuda().requires_grad_()
    b = torch.randn(256, generator=g).cuda().requires_grad_()
    xr = x.clone().requires_grad_()
    y = ops.linear_tc(xr, w, b, relu=True)
    y.backward(torch.ones_like(y))
    xd, wd, bd = (t.detach().double().requires_grad_() for t in (x, w, b))
    yd = torch.relu(xd @ wd.t() + bd)
    yd.backward(torch.ones_like(yd))
    assert float((y.detach().double() - yd.detach()).abs().max()) <= ftol * float(yd.detach().abs().max())
    same = ops.linear_forward(x, w, b, True)           # strided input, no copy
    assert torch.equal(same, y.detach())
    for got, want in ((xr.grad, xd.grad), (w.grad, wd.grad), (b.grad, bd.grad)):
        assert float((got.double() - want).abs().max()) <= gtol * float(want.abs().max())


def test_linear_on_two_input_blocks(terms):
    """fc6 / fc10: relu(fc(cat([x, x2]))) with the concatenation read in place, forward and backward."""
    from atmonr.native import ops
    ftol, gtol = (4e-6, 1e-5) if terms == 3 else (5e-5, 1e-4)
    g = torch.Generator().manual_seed(2)
    for k1, k2, n in ((256, 76, 256), (256, 24, 128), (20, 12, 8)):   # the last one is not a multiple of 8: cat fallback
        feat = torch.randn(900, k1 + 4, generator=g).cuda()
        x1 = feat[:, :k1].detach().requires_grad_()                   # a column slice, like feat[:, :hidden_dim]
        x2 = torch.randn(900, k2, generator=g).cuda().requires_grad_()
        w = (torch.randn(n, k1 + k2, generator=g) / 16).cuda().requires_grad_()
        b = torch.randn(n, generator=g).cuda().requires_grad_()
        y = ops.linear_tc(x1, w, b, relu=True, x2=x2)
        gy = torch.randn(900, n, generator=g).cuda()
        y.backward(gy)
        d = [t.detach().double().requires_grad_() for t in (x1, x2, w, b)]
        yd = torch.relu(torch.cat([d[0], d[1]], 1) @ d[2].t() + d[3])
        yd.backward(gy.double())
        assert float((y.detach().double() - yd.detach()).abs().max()) <= ftol * float(yd.detach().abs().max())
        for got, want in zip((x1.grad, x2.grad, w.grad, b.grad), (t.grad for t in d)):
            assert got.shape == want.shape
            assert float((got.double() - want).abs().max()) <= gtol * float(want.abs().max())


def test_nerf_pipeline_with_tensor_core_layers(monkeypatch, terms):
    """configs/nerf.json forward + loss + backward on the tensor-core layers against library GEMMs:
    same parameters, same draws (eval mode: no density noise; the sampler's Philox stream is keyed by
    the step counter, which both runs start from zero)."""
    import json
    from helpers import ROOT
    from atmonr.batch_loader import BatchLoader
    from atmonr.datasets.factory import get_dataset
    from atmonr.pipelines.factory import get_pipeline
    cfg = json.load(open(os.path.join(ROOT, "configs", "nerf.json")))
    ds = get_dataset(cfg["dataset"], "synthetic:H=12,W=12,seed=3")
    batch = next(iter(BatchLoader(ds, batch_size=512, shuffle=True, seed=7)))
    outs = {}
    for mode in ("lib", "tc"):
        import atmonr.models.nerf as mn
        monkeypatch.setattr(mn, "DENSE_IMPL", "tc" if mode == "tc" else "library")
        torch.manual_seed(0)
        pipe = get_pipeline(cfg["pipeline"], ds)
        pipe.send_tensors_to(0)
        pipe.eval()
        torch.manual_seed(1)
        res = pipe.forward(batch)
        loss = pipe.compute_loss(batch, res)
        loss.backward()
        grads = torch.cat([p.grad.flatten() for net in pipe.nerf.values() for p in net.parameters()])
        outs[mode] = (res["color_map_fine"].detach(), res["color_map_coarse"].detach(), float(loss), grads)
    a, b = outs["lib"], outs["tc"]
    ctol, gtol = (1e-4, 1e-3) if terms == 3 else (3e-4, 3e-3)
    for x, y in ((a[0], b[0]), (a[1], b[1])):
        assert float((x - y).abs().max()) <= ctol * float(x.abs().max())      # north_star: 1e-3 relative
    assert abs(a[2] - b[2]) <= ctol * abs(a[2])
    assert float((a[3] - b[3]).abs().max()) <= gtol * float(a[3].abs().max())
    print(f"terms={terms}: colour map rel err {float((a[0] - b[0]).abs().max() / a[0].abs().max()):.2e}, "
          f"loss rel err {abs(a[2] - b[2]) / abs(a[2]):.2e}, gradient rel err {float((a[3] - b[3]).abs().max() / a[3].abs().max()):.2e}")


def test_linear_output_mask_and_preallocated_output(terms):
    """`out_mask` (the ReLU derivative of the layer below, applied while the input gradient is written out)
    and `out` (a column block of a wider tensor receives the product): the backward chain of
    atmonr.native.nerf_mlp is built from these two."""
    from atmonr.native import ops
    g = torch.Generator().manual_seed(11)
    tol = 4e-6 if terms == 3 else 5e-5
    for m, k, n in ((700, 128, 256), (300, 256, 76), (129, 260, 256)):
        dy = torch.randn(m, k, generator=g).cuda()
        w = (torch.randn(k, n, generator=g) / k ** 0.5).cuda()          # dX = dY @ W, W (k, n)
        below = torch.randn(m, n + 3, generator=g).cuda()[:, :n]         # a column slice as the mask
        want = (dy.double() @ w.double()) * (below > 0)
        got = ops.linear_forward(dy, w, None, False, transpose=True, out_mask=below)
        assert got.shape == (m, n)
        assert float((got.double() - want).abs().max()) <= tol * float(want.abs().max())
        assert bool((got[below <= 0] == 0).all())
        wide = torch.full((m, n + 8), 7.0, device="cuda")
        ops.linear_forward(dy, w, None, False, transpose=True, out=wide)
        assert float((wide[:, :n].double() - dy.double() @ w.double()).abs().max()) <= tol * float(want.abs().max())
        assert bool((wide[:, n:] == 7.0).all())                            # columns beyond the product are untouched


def test_linear_relu_sign_bits(terms):
    """want_bits / out_bits: the forward product records the sign bits of its ReLU output, the input-gradient
    product of the layer above applies them in its epilogue; same result as masking with the activations
    themselves (out_mask), incl. a ragged last tile and a 128-wide layer."""
    from atmonr.native import ops
    g = torch.Generator().manual_seed(13)
    for m, k, n in ((1000, 76, 256), (129, 256, 128), (4096, 332, 256)):
        x = torch.randn(m, k, generator=g).cuda()
        w = (torch.randn(n, k, generator=g) / k ** 0.5).cuda()
        b = torch.randn(n, generator=g).cuda()
        y, bits = ops.linear_forward(x, w, b, True, want_bits=True)
        assert bits.shape == (m, n // 32) and bits.dtype == torch.int32
        assert torch.equal(y, ops.linear_forward(x, w, b, True))
        # the bits say exactly y > 0 (population count per row, independent of the private bit layout)
        pop = sum(((bits >> s) & 1).sum(dim=1) for s in range(32))
        assert torch.equal(pop, (y > 0).sum(dim=1))
        dy = torch.randn(m, 64, generator=g).cuda()
        w2 = (torch.randn(64, n, generator=g) / 8).cuda()
        a = ops.linear_forward(dy, w2, None, False, transpose=True, out_mask=y)
        c = ops.linear_forward(dy, w2, None, False, transpose=True, out_bits=bits)
        assert torch.equal(a, c)
        assert bool((c[y <= 0] == 0).all()) and float(c.abs().max()) > 0


def test_nerf_mlp_node_matches_the_layer_by_layer_model(terms):
    """atmonr.native.nerf_mlp.NerfMlpFn (one autograd node, hand-written backward chain) against the same
    AtmoNeRF evaluated with torch's float32 layers + autograd: outputs, parameter gradients and the gradient
    w.r.t. the encoded position (the fine pass differentiates it; the direction columns get none)."""
    import atmonr.models.nerf as mn
    torch.manual_seed(3)
    coarse, fine = mn.get_model(256, 4, [14, 14, 10], 4, False)
    g = torch.Generator().manual_seed(5)
    for net in (coarse, fine):
        net = net.cuda().eval()
        x = (torch.rand(3000, 100, generator=g) * 2 - 1).cuda()
        gc = torch.randn(3000, 4, generator=g).cuda()
        gs = torch.randn(3000, net.volume_channels, generator=g).cuda()
        outs = {}
        for impl in ("library", "tc"):
            mn.DENSE_IMPL = impl
            try:
                xr = x.clone().requires_grad_()
                net.zero_grad()
                c, s = net(xr)
                ((c * gc).sum() + (s * gs).sum()).backward()
                outs[impl] = (c.detach(), s.detach(), xr.grad.clone(), [p.grad.clone() for p in net.parameters()])
            finally:
                mn.DENSE_IMPL = "tc"
        a, b = outs["library"], outs["tc"]
        rel = lambda u, v: float((u - v).abs().max() / (u.abs().max() + 1e-30))
        rel2 = lambda u, v: float((u - v).double().norm() / (u.double().norm() + 1e-30))
        # Gradients are compared in the L2 norm at 1e-2: the two evaluations differ by ~1e-5 in the
        # pre-activations, so of the 3000 x 256 x 9 ReLU units a few dozen sit on opposite sides of zero, and
        # ONE such unit moves its row of a weight gradient by 1 / sqrt(3000) of the largest entry (measured
        # 1.8e-2 in the max norm, identical for two and three bf16 terms: it is the library product's
        # rounding against the tensor-core one, not the split).
        ftol, gtol = 1e-4, 1e-2
        errs = {"colour": rel(a[0], b[0]), "density": rel(a[1], b[1]), "dx_pos": rel2(a[2][:, :76], b[2][:, :76]),
                "params": max(rel2(pa, pb) for pa, pb in zip(a[3], b[3]))}
        print(f"terms={terms} V={net.volume_channels}:", {k: f"{v:.2e}" for k, v in errs.items()},
              [f"{n}:{rel2(pa, pb):.1e}" for (n, _), pa, pb in zip(net.named_parameters(), a[3], b[3])])
        assert errs["colour"] <= ftol and errs["density"] <= ftol, errs
        assert errs["dx_pos"] <= gtol and float(b[2][:, 76:].abs().max()) == 0, errs
        for pa, pb in zip(a[3], b[3]):
            assert pa.shape == pb.shape
        assert errs["params"] <= gtol, errs
