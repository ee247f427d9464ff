"""bench.py -- training throughput of the Instant-NGP hot path (BASELINE.json configs[1]):

    rays/s of forward + loss + backward + AdamW, 2^18-ray batches x 1024 samples per ray,
    synthetic HARP2-shaped granule (4 bands, 10/10/60/10 views), random-init weights.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched with torchrun)
    python bench.py --impl reference ...                     (CPU arm: the oracle port)

Prints ONE JSON line (rank 0). `value` = device-resident inputs; `e2e` = the same step through
the pipeline API with pinned-host batches (H2D of the batch + D2H of the loss inside the timed
region). Scaling is weak: every rank processes `--rays` rays per step.
"""

from __future__ import annotations

import argparse
import math
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "atmospheric-neural-rendering_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "training rays/sec (fwd+bwd+Adam), Instant-NGP"
UNIT = "rays/s"
OPT_CFG = {"lr": 1e-2, "betas": [0.9, 0.99], "eps": 1e-15, "weight_decay": 1e-2}
N_PARAMS = 47812560


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--rays", type=int, default=1 << 18, help="rays per step per GPU")
    ap.add_argument("--samples", type=int, default=1024)
    ap.add_argument("--granule", default="synthetic:H=256,W=256,seed=0")
    ap.add_argument("--cpu-rays", type=int, default=256, help="rays per step of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--compact-backward", action="store_true",
                    help="field backward over the samples that carry a gradient only (exact; pays off on sparse scenes)")
    ap.add_argument("--no-prefetch", action="store_true",
                    help="sample every batch in line instead of underneath the previous step's backward")
    return ap.parse_args()


def pipeline_config(samples: int) -> dict:
    cfg = json.load(open(os.path.join(ROOT, "configs", "instant_ngp.json")))
    cfg["pipeline"]["num_samples_per_ray"] = samples
    return cfg


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6), ("sw_power_cap", 7)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the same step
# ------------------------------------------------------------------------------------------
def cpu_step_rate(samples: int, rays: int, steps: int, warmup: int) -> dict:
    """Time the oracle's Instant-NGP training step (torch CPU, all host threads) on a bounded
    sample: `rays` rays x `samples` samples per step."""
    from helpers import random_params, take, tiny_scene
    from oracle.ngp import NGPOracle

    torch.set_num_threads(os.cpu_count() or 1)
    scene = tiny_scene(h=16, w=16, n_views=9)
    cfg = pipeline_config(samples)["pipeline"]
    orc = NGPOracle(cfg, scene.frame, scene.max_i, fp16=False)
    params = random_params(orc, seed=0)
    opt = orc.make_optimizer(params, OPT_CFG)
    g = torch.Generator().manual_seed(0)
    n = scene.batch["origin"].shape[0]
    times = []
    for it in range(warmup + steps):
        sel = torch.randint(0, n, (rays,), generator=g)
        batch = take(scene.batch, sel)
        u = torch.rand(rays, samples, generator=g)
        t0 = time.perf_counter()
        orc.train_step(batch, params, opt, u)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"value": rays / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} steps x {rays} rays x {samples} samples/ray, oracle/ngp.py (torch CPU fp32), "
                      f"{sec:.2f} s/step; the reference's own NGP path cannot run (tiny-cuda-nn absent)",
            "ms_per_step": sec * 1e3}


def nerf_config() -> dict:
    return json.load(open(os.path.join(ROOT, "configs", "nerf.json")))


def cpu_nerf_rate(rays: int = 1024, steps: int = 2, warmup: int = 1) -> dict:
    """the reference's pure-PyTorch NeRF path on the host cores. The
    reference tree is not present on the GPU box, so its pinned restatement (oracle/nerf.py,
    bit-checked against the reference's NeRFPipeline in tests/test_oracle_golden.py) is timed:
    forward + loss + backward + Adam, configs/nerf.json pipeline section (N_c 64, N_f 128,
    hidden 256), B = `rays`, density noise on (training mode)."""
    from helpers import take, tiny_scene
    from oracle.nerf import NeRFOracle

    torch.set_num_threads(os.cpu_count() or 1)
    scene = tiny_scene(h=16, w=16, n_views=9)
    cfg = nerf_config()["pipeline"]
    orc = NeRFOracle(cfg, scene.frame)
    params = orc.init_params(0)
    opt = torch.optim.Adam([p for m in params.values() for p in m.values()], lr=5e-4)
    g = torch.Generator().manual_seed(0)
    n = scene.batch["origin"].shape[0]
    nc, nf = cfg["sampler"]["N_c"], cfg["sampler"]["N_f"]
    times = []
    for it in range(warmup + steps):
        batch = take(scene.batch, torch.randint(0, n, (rays,), generator=g))
        u_c, u_f = torch.rand(rays, nc, generator=g), torch.rand(rays, nf, generator=g)
        noise_c = torch.randn(rays * nc, 1, generator=g)
        noise_f = torch.randn(rays * (nc + nf), cfg["num_bands"], generator=g)
        t0 = time.perf_counter()
        res = orc.forward(batch, params, u_c, u_f, noise_c, noise_f)
        loss = orc.loss(batch, res)
        opt.zero_grad()
        loss.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"value": rays / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} steps x {rays} rays x ({nc}+{nc + nf}) samples, oracle/nerf.py (torch CPU fp32), {sec:.2f} s/step"}


def gpu_nerf_rate(dataset, dev, rays: int = 4096, steps: int = 5) -> dict:
    """The native NeRF pipeline (configs/nerf.json) on this GPU: rays/s of fwd + loss + bwd + Adam."""
    from atmonr.batch_loader import BatchLoader
    from atmonr.pipelines.factory import get_pipeline

    cfg = nerf_config()
    pipe = get_pipeline(cfg["pipeline"], dataset)
    pipe.send_tensors_to(dev.index)
    opt = pipe.get_optimizer(cfg["trainer"]["optimizer"])
    batch = next(iter(BatchLoader(dataset, batch_size=rays, shuffle=True, seed=7)))

    def step():
        loss = pipe.compute_loss(batch, pipe.forward(batch))
        opt.zero_grad()
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": rays * 1e3 / ms, "unit": UNIT, "rays_per_step": rays, "ms_per_step": ms,
            "note": "configs/nerf.json, coarse 64 + fine 192 samples, hidden 256; MLP layers are "
                    + ("tcgen05 bf16x3-split products (csrc/linear_tc.cu)" if _nerf_dense_impl() == "tc"
                       else "cuBLAS fp32 GEMMs (cross-check)")}


def _nerf_dense_impl() -> str:
    import atmonr.models.nerf as mn
    return mn.DENSE_IMPL


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
    cb = cpu_step_rate(args.samples, args.cpu_rays, steps, warm)
    line = {
        "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": f"Instant-NGP train step, bounded CPU sample of configs[1]: {args.cpu_rays} rays x "
                               f"{args.samples} samples/ray per step", "l2": "n/a (CPU)"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------
def run_native(args) -> None:
    from atmonr import distributed as dist
    from atmonr.batch_loader import BatchLoader
    from atmonr.datasets.factory import get_dataset
    from atmonr.native import lib as L
    from atmonr.pipelines.factory import get_pipeline

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device; there is no CPU fallback")
    rank, world, local = dist.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L.load()

    cfg = pipeline_config(args.samples)
    torch.manual_seed(0)
    dataset = get_dataset(cfg["dataset"], args.granule)
    pipe = get_pipeline(cfg["pipeline"], dataset)
    pipe.send_tensors_to(local)
    dist.broadcast_parameters(pipe.parameters())
    opt = pipe.get_optimizer(OPT_CFG)
    B, K, W = args.rays, args.steps, args.warmup

    # fixed set of batches, each rank its own rays (weak scaling)
    loader = BatchLoader(dataset, batch_size=B, shuffle=True, seed=1234 + rank)
    keys = ("origin", "dir", "len", "rad", "irgb_idx")
    batches = []
    for b in loader:
        if b["origin"].shape[0] == B:
            batches.append({k: b[k].contiguous() for k in keys})
        if len(batches) >= 4:
            break
    assert batches, "granule too small for the requested batch size"
    host = [{k: v.cpu().pin_memory() for k, v in b.items()} for b in batches]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    prefetch = not args.no_prefetch
    from atmonr.native import fused as _fused
    if args.compact_backward:
        _fused.COMPACT_BWD = True

    def step(batch, upcoming=None):
        # the NEXT batch is announced first: its sample points are computed on a side stream
        # underneath this step's backward (InstantNGPPipeline.prefetch); one sampler launch per step
        if prefetch and upcoming is not None:
            pipe.prefetch(upcoming)
        res = pipe.forward(batch)
        loss = pipe.compute_loss(batch, res)
        opt.zero_grad()
        loss.backward()
        dist.all_reduce_gradients(opt)
        opt.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item())

    # ---- settle (allocator pools, lazy module loading: the first ~6 steps of a process run up to
    # 1.6x slower), then the W warm-up steps, then the device-resident measurement ----
    nb = len(batches)
    for i in range(4):
        step(batches[i % nb], batches[(i + 1) % nb])
    for i in range(W):
        step(batches[i % nb], batches[(i + 1) % nb])
    clocks = ClockSampler(local)
    if rank == 0 and os.environ.get("ATMONR_BENCH_NO_CLOCKS") != "1":
        clocks.start()
    L.STATS = L.CallStats(timed=False)
    profile_range = os.environ.get("ATMONR_CUDA_PROFILER_RANGE") == "1"  # ncu --profile-from-start off
    if profile_range:
        torch.cuda.profiler.start()
    ms_total = timed(K, lambda i: step(batches[i % nb], batches[(i + 1) % nb]))
    if profile_range:
        torch.cuda.profiler.stop()
    launches = L.STATS.launches
    L.STATS = None
    n_act = pipe.fused_state.last.get("n_active") if pipe.fused_state is not None else None
    active_fraction = float(n_act.item()) / (B * args.samples) if n_act is not None else 1.0
    clock_info = clocks.stop() if rank == 0 else {}
    ms_step = ms_total / K
    value = world * B * 1e3 / ms_step

    # ---- per-kernel durations (CUDA events around every C-ABI call, separate pass) ----
    # (the sampler runs in line here, not prefetched on its side stream, so that every kernel is timed alone)
    L.STATS = L.CallStats(timed=True)
    for i in range(min(K, 3)):
        step(batches[i % nb], None)
    dur = L.STATS.durations_ms()
    calls = dict(L.STATS.calls)
    L.STATS = None
    per_step = {k: sum(v) / min(K, 3) for k, v in dur.items()}
    top = max(per_step, key=per_step.get)
    M = B * args.samples
    # SURVEY 8d per-unit figures: 512 B gathered (fwd) / 512 B scattered (bwd) per sample -- the
    # backward reads the forward's cached features (64 B/sample) instead of re-gathering the table;
    # AdamW 28 B + 2 B fp16 shadow per parameter.
    alg_bytes = {
        "atmonr_ngp_field_fwd": 512 * M, "atmonr_ngp_field_bwd": 1024 * M,
        "atmonr_ngp_field_fwd_tc": 512 * M, "atmonr_ngp_field_bwd_tc": 512 * M,
        # the compact backward scatters for the listed samples only: count those
        "atmonr_ngp_field_bwd_tc_compact": int(512 * M * active_fraction),
        "atmonr_composite_bwd_compact": int((28 + 24 * active_fraction) * M),
        "atmonr_adamw_step": 30 * N_PARAMS, "atmonr_ngp_sample_points": 16 * M + 28 * B,
        "atmonr_composite_fwd": 24 * M, "atmonr_composite_bwd": 44 * M,
    }
    traffic_file = os.path.join(ROOT, "profiles", "ncu_dram_traffic.json")
    traffic = json.load(open(traffic_file)) if os.path.exists(traffic_file) else {}
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    n_top = max(1, calls.get(top, 1) // min(K, 3))
    achieved = alg_bytes.get(top, 0) / (per_step[top] / n_top * 1e-3) / 1e9 if top in alg_bytes else None
    roofline = {
        "kernel": top, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
        "frac": (achieved / hbm_peak) if achieved else None,
        "traffic": traffic.get(top, {}).get("dram_bytes_per_launch") if traffic.get("rays") == B else None,
        "algorithmic_bytes_per_launch": alg_bytes.get(top),
        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
        "note": "algorithmic bytes = table gathers / gradient scatters per SURVEY 8d (each 2-feature RED counted as "
                "4 B like the reference's half2 atomics); this traffic is served by L2 (the table and most of its "
                "gradient are L2-resident), so frac is table traffic quoted against the HBM copy peak",
        "achieved_gbs_by_kernel": {k: round(alg_bytes[k] / (per_step[k] / max(1, calls.get(k, 1) // min(K, 3)) * 1e-3) / 1e9, 1)
                                   for k in per_step if k in alg_bytes},
        "ms_per_step_by_kernel": {k: round(v, 3) for k, v in sorted(per_step.items(), key=lambda kv: -kv[1])},
    }

    # ---- end to end through the pipeline API with host buffers ----
    # Every step copies ONE batch from pinned host memory (the next step's, double-buffered like the
    # sample points) and reads the loss back.
    staged = {}

    def h2d(i):
        return {k: v.to(dev, non_blocking=True) for k, v in host[i % len(host)].items()}

    def e2e_step(i):
        batch = staged.pop(i, None) or h2d(i)
        if prefetch:
            staged[i + 1] = h2d(i + 1)
        return step(batch, staged.get(i + 1)).item()

    e2e_step(-1)
    staged.clear()
    staged[0] = h2d(0)
    if prefetch:
        pipe.prefetch(staged[0])
    ms_e2e = timed(K, e2e_step) / K
    staged.clear()
    e2e = {"value": world * B * 1e3 / ms_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
           "ms_per_step": ms_e2e}

    # ---- extraction: voxel queries/s (BASELINE.json metric, second half), inputs resident in HBM ----
    # SURVEY 8d workload: a dense grid of voxel columns over the granule x 81 altitudes (0..20 km, step
    # 250 m), float64 ECEF points normalised like scripts/extract.py:81, one call = 32768 columns x 81.
    # Each rank queries its own slab of columns (no communication).
    from atmonr.datasets.harp2_extract import HARP2VoxelGridExtractDataset
    n_vox = 32768 * 81
    lat_span_m = math.radians(float(dataset.lat[~dataset.lat.isnan()].max() - dataset.lat[~dataset.lat.isnan()].min())) * 6378137.0
    h_step = lat_span_m / 620.0   # ~620 x ~500 columns over the 5 x 5 degree granule: > 4 x 32768 columns
    grid_ds = HARP2VoxelGridExtractDataset(dataset, horizontal_step=h_step, alt_step=250.0)
    n_alt = int(grid_ds.sample_alt.shape[0])
    n_cols = len(grid_ds) // n_alt
    per_call = min(32768, n_cols // max(world, 1))
    lo = (rank * per_call) % max(n_cols - per_call + 1, 1)
    n_vox = per_call * n_alt
    vox = ((grid_ds.xyz[lo * n_alt:(lo + per_call) * n_alt].to(dev) - dataset.offset.to(dev)) / dataset.scale).contiguous()
    assert vox.dtype == torch.float64 and vox.shape == (n_vox, 3)
    pipe.eval()
    with torch.no_grad():
        for _ in range(2):
            pipe.extract(vox)
        ms_ext = timed(5, lambda i: pipe.extract(vox)) / 5
    pipe.train()
    extract = {"value": world * n_vox * 1e3 / ms_ext, "unit": "voxels/s", "voxels_per_call_per_gpu": n_vox,
               "ms_per_call": ms_ext,
               "workload": f"voxel grid, {per_call} columns x {n_alt} altitudes per call, float64 points (scripts/extract.py voxelgrid mode)"}

    if rank != 0:
        return
    # the NeRF line and the CPU baselines belong to the N = 1 run (rank 0 would otherwise keep the
    # other ranks' GPUs idle for ~20 s of host work in every run of the scaling sweep)
    solo = world == 1
    nerf = gpu_nerf_rate(dataset, dev) if solo else None
    cpu = cpu_step_rate(args.samples, args.cpu_rays, 2, 1) if solo and not args.no_cpu_baseline else None
    cpu_nerf = cpu_nerf_rate() if solo and not args.no_cpu_baseline else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 table/MLP operands, f32 accumulate, f64 geodesy", "data": "synthetic",
        "config": {
            "workload": f"Instant-NGP (configs/instant_ngp.json) train step, {B} rays/GPU x {args.samples} samples/ray, "
                        f"{args.granule} HARP2-shaped granule, 4 bands 10/10/60/10 views",
            "rays_per_gpu": B, "samples_per_ray": args.samples, "parallelism": f"dp{world}",
            "backward": (f"samples with a non-zero incoming gradient only ({active_fraction:.3f} of all samples in the last step; exact)"
                         if _fused.COMPACT_BWD else "dense (every sample)"),
            "l2": "inputs larger than L2: per-step working set (x01, sigma, colour, gradients) is several GB",
        },
        "clocks": clock_info, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "extract": extract,
        "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None,
        "nerf": nerf, "cpu_baseline_nerf": cpu_nerf,
    }
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
