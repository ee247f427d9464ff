"""The AtmoNeRF MLP (reference: src/atmonr/models/nerf.py:48-93) as ONE autograd node over the tcgen05
dense-layer kernels (csrc/linear_tc.cu).

Forward: the eleven layers in order, bias + ReLU in each product's epilogue, the two concatenations
(fc6: [h5 | x_pos], fc10: [feat | x_dir]) read in place. Backward: the chain written out by hand, so that

  * every input-gradient product writes the PRE-activation gradient of the layer below it (the ReLU
    derivative of that layer's output is applied in the product's epilogue): no product reads a mask in
    its main loop (measured round 2: with the mask staged next to the gradient operand, half of the warp
    samples of the weight-gradient kernel sat on that load). The derivative travels as SIGN BITS written
    by the forward product's epilogue (32 bytes per 256-wide row), not as a second read of the activations;
  * columns nobody differentiates are never computed (the 24 direction columns of fc10's input gradient;
    the position-encoding gradient unless the sample distances carry one, i.e. in the coarse pass);
  * the density head's gradient is dropped into the last columns of fc9's output gradient by the product
    that writes the first 256 (no concatenation, no per-layer autograd bookkeeping, no slice copies).

The parameters stay the `fc1`..`fc11` nn.Linear modules of atmonr.models.nerf.AtmoNeRF (checkpoints
interchange with the reference); this node only borrows their tensors.
"""

from __future__ import annotations

import torch

from atmonr.native import ops

_f32 = torch.float32


class NerfMlpFn(torch.autograd.Function):
    """(x (M, pos + dir), density_noise (M, V) | None, w1, b1, ..., w11, b11) -> (rgb_raw (M, out), sigma_pre (M, V)).
    rgb_raw is fc11's output (before the sigmoid of models/nerf.py:91), sigma_pre the density head of fc9 plus
    the noise (before the ReLU of :71)."""

    @staticmethod
    def forward(ctx, x, noise, pos_channels, hidden, *params):
        w = params[0::2]
        b = params[1::2]
        x, _ = ops._rows(x)
        x_pos, x_dir = x[:, :pos_channels], x[:, pos_channels:]
        acts, bits = [], []            # post-ReLU outputs of fc1..fc8 (inputs of fc2..fc9) and their sign bits
        use_bits = hidden % 128 == 0   # the sign-bit form of the epilogue covers whole 128-column halves

        def layer(xin, k, x2=None):
            if use_bits:
                h, sb = ops.linear_forward(xin, w[k], b[k], True, x2=x2, want_bits=True)
            else:
                h, sb = ops.linear_forward(xin, w[k], b[k], True, x2=x2), None
            acts.append(h)
            bits.append(sb)
            return h

        h = layer(x_pos, 0)
        for k in range(1, 5):          # fc2..fc5
            h = layer(h, k)
        h = layer(h, 5, x2=x_pos)      # fc6: skip connection
        for k in (6, 7):               # fc7, fc8
            h = layer(h, k)
        # fc9: (M, hidden + V), no activation; rows padded to a multiple of 4 floats so that fc10 and the weight
        # gradients read them with 16-byte loads (hidden + 1 columns in the coarse network)
        n9 = w[8].shape[0]
        feat = ops.linear_forward(h, w[8], b[8], False,
                                  out=torch.empty((x.shape[0], (n9 + 3) // 4 * 4), device=x.device, dtype=_f32))[:, :n9]
        hid_bits = None
        if w[9].shape[0] % 128 == 0:
            hid, hid_bits = ops.linear_forward(feat[:, :hidden], w[9], b[9], True, x2=x_dir, want_bits=True)   # fc10
        else:
            hid = ops.linear_forward(feat[:, :hidden], w[9], b[9], True, x2=x_dir)
        rgb = ops.linear_forward(hid, w[10], b[10], False)         # fc11
        sigma = feat[:, hidden:]
        sigma = sigma + noise if noise is not None else sigma.clone()
        ctx.save_for_backward(x, feat, hid, *acts, *w)
        ctx.bits, ctx.hid_bits = bits, hid_bits
        ctx.pos_channels, ctx.hidden = pos_channels, hidden
        ctx.x_needs_grad = ctx.needs_input_grad[0]
        return rgb, sigma

    @staticmethod
    def backward(ctx, d_rgb, d_sigma):
        saved = ctx.saved_tensors
        x, feat, hid = saved[:3]
        acts = saved[3:11]
        w = saved[11:]
        pc, hd = ctx.pos_channels, ctx.hidden
        x_pos, x_dir = x[:, :pc], x[:, pc:]
        m = x.shape[0]
        grads = [None] * 22
        bits, hid_bits = ctx.bits, ctx.hid_bits

        def dinput(dy, weight, below, below_bits):
            """dy @ weight, zeroed where the layer below did not fire (its sign bits, or its activations)."""
            if below_bits is not None:
                return ops.linear_forward(dy, weight, None, False, transpose=True, out_bits=below_bits)
            return ops.linear_forward(dy, weight, None, False, transpose=True, out_mask=below)

        def wgrad(k, dy, xin, x2=None):
            grads[2 * k], grads[2 * k + 1] = ops.linear_weight_grad(dy, xin, x2=x2, want_bias=True)

        d_rgb = ops._c(d_rgb, _f32)
        # fc11: pre-activation gradient of fc10 comes out masked by hid > 0
        wgrad(10, d_rgb, hid)
        d_hid = dinput(d_rgb, w[10], hid, hid_bits)
        # fc10: only the `feat` columns of its input carry a gradient; they land in the first `hidden`
        # columns of fc9's output gradient, the density gradient goes into the remaining ones
        wgrad(9, d_hid, feat[:, :hd], x2=x_dir)
        v = feat.shape[1] - hd
        ld9 = (feat.shape[1] + 3) // 4 * 4
        d_feat = torch.empty((m, ld9), device=x.device, dtype=_f32)
        ops.linear_forward(d_hid, w[9][:, :hd], None, False, transpose=True, out=d_feat)
        if d_sigma is not None:
            d_feat[:, hd:hd + v] = d_sigma
        else:
            d_feat[:, hd:hd + v] = 0
        d9 = d_feat[:, :hd + v]
        # fc9 .. fc7
        wgrad(8, d9, acts[7])
        d = dinput(d9, w[8], acts[7], bits[7])
        wgrad(7, d, acts[6])
        d = dinput(d, w[7], acts[6], bits[6])
        wgrad(6, d, acts[5])
        d = dinput(d, w[6], acts[5], bits[5])
        # fc6: input [h5 | x_pos]
        wgrad(5, d, acts[4], x2=x_pos)
        d_pos = None
        if ctx.x_needs_grad:
            d_pos = ops.linear_forward(d, w[5][:, hd:], None, False, transpose=True)
        d = dinput(d, w[5][:, :hd], acts[4], bits[4])
        # fc5 .. fc2
        for k in (4, 3, 2, 1):
            wgrad(k, d, acts[k - 1])
            d = dinput(d, w[k], acts[k - 1], bits[k - 1])
        # fc1
        wgrad(0, d, x_pos)
        dx = None
        if ctx.x_needs_grad:
            # the direction columns carry no gradient (the rays are data); the position columns are written
            # in place by the product (row stride = the width of x)
            dx = torch.zeros((m, x.shape[1]), device=x.device, dtype=_f32)
            ops.linear_forward(d, w[0], None, False, transpose=True, out=dx)
            dx[:, :pc] += d_pos
        return (dx, None, None, None, *grads)
