"""Diagnostic for the field backward at the BENCHMARK configuration ("78 % zero rows").

    python scripts/diag_backward_rows.py [--rays 262144] [--samples 1024]

On the bench workload (bench.py: synthetic granule, random-init parameters, Philox draws) it reports

  1. the distribution of the incoming per-sample gradients (dL/d sigma_raw, dL/d colour_raw) relative
     to their maximum, i.e. what the power-of-two operand scale of k_field_bwd_tc2 has to span;
  2. the fraction of sample rows whose dL/d(encoded features) is EXACTLY zero in float32 arithmetic
     (torch float32 back-propagation through the two MLPs with the kernels' fp16 weights), and the
     fraction that additionally becomes zero when the gradient operands are rounded to fp16 under
     the kernel's scale (emulated in torch), with the share of the gradient's L2 norm those rows carry;
  3. || g_tc - g_simt ||_2 / || g_simt ||_2 for the hash-table gradient and the MLP weight gradients
     (tcgen05 path with fp16 gradient operands vs the float32 SIMT cross-check kernels), and the colour maps.

Prints one JSON document; profiles/r3_backward_rows.json is a committed copy of a B200 run.
"""

from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atmospheric-neural-rendering_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def l2_rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-300))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=1 << 18)
    ap.add_argument("--samples", type=int, default=1024)
    ap.add_argument("--emulate-rays", type=int, default=4096, help="rays of the torch float32 emulation (part 2)")
    ap.add_argument("--granule", default="synthetic:H=256,W=256,seed=0")
    args = ap.parse_args()

    import bench
    from atmonr.batch_loader import BatchLoader
    from atmonr.datasets.factory import get_dataset
    from atmonr.native import fused, lib as L, ops
    from atmonr.pipelines.factory import get_pipeline

    assert torch.cuda.is_available(), "needs a CUDA device"
    torch.cuda.set_device(0)
    L.load()
    cfg = bench.pipeline_config(args.samples)
    torch.manual_seed(0)
    ds = get_dataset(cfg["dataset"], args.granule)
    pipe = get_pipeline(cfg["pipeline"], ds)
    pipe.send_tensors_to(0)
    st = pipe.fused_state
    B, N = args.rays, args.samples
    batch = next(iter(BatchLoader(ds, batch_size=B, shuffle=True, seed=1234)))
    out: dict = {"rays": B, "samples_per_ray": N}

    # ---- 3. tc vs simt on identical draws (the draw counter is reset before each run) ----
    grads, maps = {}, {}
    for impl in ("tc", "simt"):
        fused.FIELD_IMPL = impl
        st.step = 0
        for p in pipe.parameters():
            p.grad = None
        res = pipe.forward(batch)
        loss = pipe.compute_loss(batch, res)
        loss.backward()
        torch.cuda.synchronize()
        grads[impl] = {n: getattr(pipe, n).params.grad.clone() for n in ("pos_encoder", "pos_mlp", "dir_mlp", "surf_encoder", "surf_mlp")}
        maps[impl] = (res["color_map_fine"].detach().clone(), float(loss))
    fused.FIELD_IMPL = "tc"
    out["tc_vs_simt"] = {
        "loss_tc": maps["tc"][1], "loss_simt": maps["simt"][1],
        "color_map_l2_rel": l2_rel(maps["tc"][0], maps["simt"][0]),
        "color_map_max_rel": float((maps["tc"][0] - maps["simt"][0]).abs().max() / maps["simt"][0].abs().max()),
        "grad_l2_rel": {n: l2_rel(grads["tc"][n], grads["simt"][n]) for n in grads["tc"]},
        "grad_max_rel": {n: float((grads["tc"][n] - grads["simt"][n]).abs().max() / grads["simt"][n].abs().max()) for n in grads["tc"]},
        "table_grad_nonzero_entries": {k: int((grads[k]["pos_encoder"] != 0).sum()) for k in grads},
    }
    del grads

    # ---- 1. incoming gradients of the field backward (recomputed from the saved forward buffers) ----
    last = st.last
    z, sig, col, cs = last["z"], last["sigma_raw"], last["color_raw"], last["color_surf_raw"]
    cmap, catmo, csurf, tsurf, _, _ = ops.composite_forward(z, col, sig, cs, st.z_scale, relu=True, want_weights=False, want_alpha=False)
    cm = cmap.detach().requires_grad_()
    ops.band_loss(cm, batch["irgb_idx"], batch["rad"], pipe.max_i, pipe.loss_name).backward()
    absmax = torch.zeros(1, device=z.device)
    dcolor, dsigma, _ = ops.composite_backward(z, col, sig, cs, catmo, tsurf, cm.grad, cm.grad, st.z_scale, relu=True, grad_absmax=absmax)
    amax = float(absmax)
    S = 2.0 ** max(-60.0, min(60.0, float(torch.floor(torch.log2(torch.tensor(2048.0 / amax))))))
    out["incoming"] = {"absmax": amax, "operand_scale_log2": float(torch.log2(torch.tensor(S)))}
    rows_zero_in = (dsigma.view(-1) == 0) & (dcolor.view(-1, 4) == 0).all(1)
    out["incoming"]["rows_with_zero_incoming_gradient"] = float(rows_zero_in.float().mean())
    out["incoming"]["sigma_raw_positive"] = float((sig > 0).float().mean())
    for name, t in (("dsigma", dsigma.view(-1)), ("dcolor", dcolor.view(-1))):
        a = t.abs()
        nz = a[a > 0]
        q = torch.tensor([0.001, 0.01, 0.1, 0.5, 0.9, 0.99, 1.0], device=a.device)
        # quantiles on a strided subsample (torch.quantile is limited to 16M elements)
        sub = nz[:: max(1, nz.numel() // 8_000_000)]
        out["incoming"][name] = {
            "nonzero_fraction": float(nz.numel() / a.numel()),
            "log2_of_value_over_absmax_quantiles": dict(zip(["q0.001", "q0.01", "q0.1", "q0.5", "q0.9", "q0.99", "max"],
                                                            [round(float(v), 2) for v in torch.log2(torch.quantile(sub.float(), q) / amax)])),
            "fraction_below_fp16_subnormal_after_scale": float((nz * S < 2.0 ** -25).float().sum() / a.numel()),
            "fraction_in_fp16_subnormal_range_after_scale": float(((nz * S >= 2.0 ** -25) & (nz * S < 2.0 ** -14)).float().sum() / a.numel()),
        }

    # ---- 2. float32 emulation of the MLP backward on the first rays: which rows are zero, and why ----
    R = min(args.emulate_rays, B)
    M = R * N
    x01 = last["x01"][:M]
    enc = ops.HashGridFn.apply(x01, pipe.pos_encoder.params.detach(), pipe.pos_encoder.table_f16(), st.grid3).detach()
    pw = pipe.pos_mlp.weights_f16().float()
    dw = pipe.dir_mlp.weights_f16().float()
    W1, W2 = pw[:1024].view(32, 32), pw[1024:1536].view(16, 32)
    D1, D2, D3 = dw[:1024].view(32, 32), dw[1024:2048].view(32, 32), dw[2048:2560].view(16, 32)
    dirs = batch["dir"][:R].float()
    v = dirs * 2 - 1
    sh = torch.stack([torch.full_like(v[:, 0], 0.28209479177387814), -0.48860251190291987 * v[:, 1],
                      0.48860251190291987 * v[:, 2], -0.48860251190291987 * v[:, 0]], 1).half().float()
    sh = sh[:, None].expand(R, N, 4).reshape(M, 4)
    ds_in, dc_in = dsigma.view(-1)[:M], dcolor.view(-1, 4)[:M]

    def backprop(fp16_operands: bool):
        q = (lambda t: (t * S).half().float() / S) if fp16_operands else (lambda t: t)
        h16 = lambda t: t.half().float()
        h = h16(torch.relu(enc @ W1.t()))
        po = h16(h @ W2.t())
        din = torch.cat([sh, po[:, 1:16], torch.ones(M, 13, device=enc.device)], 1)
        h1 = h16(torch.relu(din @ D1.t()))
        h2 = h16(torch.relu(h1 @ D2.t()))
        dout = torch.zeros(M, 16, device=enc.device)
        dout[:, :4] = dc_in
        dh2 = q((q(dout) @ D3) * (h2 > 0))
        dh1 = q((dh2 @ D2) * (h1 > 0))
        ddin = dh1 @ D1
        dpo = torch.cat([ds_in[:, None], ddin[:, 4:19]], 1)
        dh = q((q(dpo) @ W2) * (h > 0))
        return dh @ W1                                        # dL/d(encoded features), (M, 32)

    g32 = backprop(False)
    g16 = backprop(True)
    zero32 = (g32 == 0).all(1)
    zero16 = (g16 == 0).all(1)
    # the scatter works per (row, level): a level's pair of features
    pair32 = (g32.view(M, 16, 2) == 0).all(2)
    pair16 = (g16.view(M, 16, 2) == 0).all(2)
    e32 = float(g32.double().pow(2).sum())
    lost = zero16 & ~zero32
    out["feature_gradient_rows"] = {
        "rays_emulated": R,
        "zero_rows_float32": float(zero32.float().mean()),
        "zero_rows_fp16_operands": float(zero16.float().mean()),
        "zero_row_level_pairs_float32": float(pair32.float().mean()),
        "zero_row_level_pairs_fp16_operands": float(pair16.float().mean()),
        "l2_share_of_rows_lost_to_fp16": float(g32[lost].double().pow(2).sum() / e32) ** 0.5 if e32 > 0 else 0.0,
        "l2_rel_fp16_vs_float32": l2_rel(g16, g32),
        "zero_rows_float32_given_sigma_raw_nonpositive": float(zero32[(sig.view(-1)[:M] <= 0)].float().mean()),
        "zero_rows_float32_given_sigma_raw_positive": float(zero32[(sig.view(-1)[:M] > 0)].float().mean()),
        "dead_pos_hidden_units_mean": float((torch.relu(enc @ W1.t()) <= 0).float().mean()),
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
