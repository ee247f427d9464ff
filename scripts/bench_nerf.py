"""NeRF training step (configs/nerf.json: coarse 64 + fine 192 samples, hidden 256) on cuda:0:
rays/s of forward + loss + backward + Adam, CUDA-event timed. The dense layers run on tcgen05
(csrc/linear_tc.cu); `--library` times torch's float32 GEMMs instead (cross-check). Prints one JSON line."""

from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "atmospheric-neural-rendering_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def main() -> None:
    import bench
    from atmonr.datasets.factory import get_dataset
    from atmonr.native import lib as L

    if not torch.cuda.is_available():
        raise SystemExit("bench_nerf.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(0)
    L.load()
    import atmonr.models.nerf as mn
    if "--library" in sys.argv:
        mn.DENSE_IMPL = "library"
    rays = int(os.environ.get("ATMONR_NERF_RAYS", "4096"))
    ds = get_dataset(bench.nerf_config()["dataset"], "synthetic:H=64,W=64,seed=0")
    out = bench.gpu_nerf_rate(ds, torch.device("cuda", 0), rays=rays, steps=int(os.environ.get("ATMONR_NERF_STEPS", "10")))
    from atmonr.native import ops
    out["dense_layers"] = f"tcgen05, {ops.LINEAR_TERMS} bf16 terms per operand (atmonr_linear_fwd_tc / _dw_tc)" if mn.DENSE_IMPL == "tc" \
        else "library float32 GEMMs"
    flop = 922e6 * rays          # 922 MFLOP per ray, forward + backward
    out["tflops_fp32_equivalent"] = flop / (out["ms_per_step"] * 1e-3) / 1e12
    print(json.dumps(out))


if __name__ == "__main__":
    main()
