"""bench.py -- training throughput of the Instant-NGP hot path (BASELINE.json configs[1]):

    rays/s of forward + loss + backward + AdamW, 2^18-ray batches x 1024 samples per ray,
    synthetic HARP2-shaped granule (4 bands, 10/10/60/10 views), random-init weights.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched with torchrun)
    python bench.py --impl reference ...                     (CPU arm: the oracle port)

Prints ONE JSON line (rank 0). `value` = device-resident inputs; `e2e` = the same step through
the pipeline API with pinned-host batches (H2D of the batch + D2H of the loss inside the timed
region). Scaling is weak: every rank processes `--rays` rays per step; for N > 1 the line also carries
`strong_scaling` (the SAME global batches of `--rays` and of 8192 rays split over the N ranks).
The loss of every timed step is checked to be finite and identical on all ranks' reductions.
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
    ap.add_argument("--granule", default="synthetic:H=512,W=512,seed=0", help="SURVEY 8d bench granule: 512 x 512 pixels x 90 views")
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

    from atmonr import distributed as dist
    world, rank = dist.world_size(), dist.rank()
    cfg = nerf_config()
    torch.manual_seed(0)
    pipe = get_pipeline(cfg["pipeline"], dataset)
    pipe.send_tensors_to(dev.index)
    dist.broadcast_parameters(pipe.parameters())
    opt = pipe.get_optimizer(cfg["trainer"]["optimizer"])
    # weak scaling (BASELINE.json configs[4]: NeRF on 8 GPUs): `rays` rays per GPU, each rank its own rays,
    # gradients of the two MLPs (1.2 M parameters) all-reduced over NCCL
    batch = next(iter(BatchLoader(dataset, batch_size=rays, shuffle=True, seed=7 + rank)))
    last = []

    def step():
        loss = pipe.compute_loss(batch, pipe.forward(batch))
        opt.zero_grad()
        loss.backward()
        dist.all_reduce_gradients(opt)
        opt.step()
        last[:] = [loss.detach()]

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    for _ in range(3):
        step()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t)
    assert bool(torch.isfinite(last[0])), "NeRF step produced a non-finite loss"
    peaks = _load_json("MEASURED_PEAKS.json")
    tpeak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    flop32 = 922e6 * rays                      # 922 MFLOP per ray, forward + backward, float32-equivalent
    from atmonr.native import ops as _ops
    terms = _ops.LINEAR_TERMS
    products = {3: 6, 2: 3}[terms] if _nerf_dense_impl() == "tc" else 1
    tf = flop32 * products / (ms * 1e-3) / 1e12    # per GPU
    c = _load_json("profiles", "ncu_counters.json")
    roof = {"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
            "float32_equivalent_TFLOPs": flop32 / (ms * 1e-3) / 1e12, "tensor_products_per_float32_product": products,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400 TFLOP/s",
            "tensor_pipe_pct_ncu": {k: c[k].get("tensor_pipe_pct") for k in ("atmonr_linear_fwd_tc", "atmonr_linear_dw_tc") if k in c},
            "traffic": None,
            "note": "whole NeRF step (MLP + sampling + compositing + Adam) against the dense bf16 tensor peak: each float32 "
                    f"product is {products} bf16 tensor-core products ({terms}-term split of both operands), counted as executed"}
    return {"value": world * rays * 1e3 / ms, "unit": UNIT, "rays_per_step_per_gpu": rays, "n_gpus": world, "ms_per_step": ms,
            "roofline": roof,
            "note": "configs/nerf.json, coarse 64 + fine 192 samples, hidden 256; MLP layers are "
                    + (f"tcgen05 products of {terms}-term bf16 splits (csrc/linear_tc.cu)" if _nerf_dense_impl() == "tc"
                       else "cuBLAS fp32 GEMMs (cross-check)")}


def _load_json(*parts) -> dict:
    path = os.path.join(ROOT, *parts)
    try:
        return json.load(open(path))
    except (OSError, ValueError):
        return {}


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
    dp_sharded = world > 1 and os.environ.get("ATMONR_DP_SHARD", "1") != "0"
    if dp_sharded:
        opt.shard_large_parameters()      # tables: reduce-scatter -> AdamW on 1/world -> all-gather of the fp16 shadow
    B, K, W = args.rays, args.steps, args.warmup
    if os.environ.get("ATMONR_L2_PERSIST"):   # tuning experiment: pin the fp16 table in the L2's persisting set-aside
        t16 = pipe.pos_encoder.table_f16()
        L.call("atmonr_l2_persist", L.ptr(t16), t16.numel() * 2, float(os.environ["ATMONR_L2_PERSIST"]), L.stream())

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

    losses = []   # device scalars of the timed steps (read back after the timed region)

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
        losses.append(loss.detach())
        return loss

    def check_losses(what):
        """Every timed step produced a finite loss; the parameters stay identical on all ranks (same
        summed gradients, same update): the checksum of the MLP weights is compared across ranks."""
        vals = torch.stack(losses).float()
        losses.clear()
        assert bool(torch.isfinite(vals).all()), f"{what}: non-finite loss {vals.tolist()}"
        if world > 1:
            chk = torch.stack([p.detach().double().sum() for p in (pipe.pos_mlp.params, pipe.dir_mlp.params, pipe.surf_mlp.params)])
            lo, hi = chk.clone(), chk.clone()
            torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
            torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
            assert torch.equal(lo, hi), f"{what}: parameters differ between ranks"
        return [float(v) for v in vals]

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
    losses.clear()
    ms_total = timed(K, lambda i: step(batches[i % nb], batches[(i + 1) % nb]))
    step_losses = check_losses("device-resident run")
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
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tensor_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s / 1400 TFLOP/s (B200_PROFILING.md)"
    counters = _load_json("profiles", "ncu_counters.json")       # per launch, one ncu --set full capture (summarise_ncu.py)
    l2u = _load_json("profiles", "r3_l2_ubench.json").get("patterns", {})   # measured L2 gather / RED peaks (profiles/ubench)
    l2_read_peak = l2u.get("gather4", {}).get("sector_GBps")     # 32-byte sectors/s x 32 B, independent random gathers
    l2_red_peak = l2u.get("red8", {}).get("sector_GBps")         # RED sector requests/s x 32 B
    n_per = {k: max(1, calls.get(k, 1) // min(K, 3)) for k in per_step}

    def kernel_detail(k):
        """Three views of one kernel from the SAME run's CUDA-event time: algorithmic bytes vs the HBM copy
        peak (the contract's recipe), executed L2 sectors vs the measured L2 gather / RED peaks, DRAM bytes
        vs the HBM peak; the sector / byte counts and the tensor-pipe share come from the committed ncu capture
        of the same kernel at the same ray count (counts do not depend on the clock; durations are live)."""
        sec = per_step[k] / n_per[k] * 1e-3
        d = {"ms": round(per_step[k] / n_per[k], 3)}
        if k in alg_bytes:
            d["algorithmic_GBps"] = round(alg_bytes[k] / sec / 1e9, 1)
            d["algorithmic_frac_of_hbm_peak"] = round(alg_bytes[k] / sec / 1e9 / hbm_peak, 3)
        c = counters.get(k)
        if c and c.get("workload_rays") == B:
            rd, red = c.get("l2_sectors_read") or 0, c.get("l2_sectors_red") or 0
            d["executed_l2_read_GBps"] = round(rd * 32 / sec / 1e9, 1)
            d["executed_l2_red_GBps"] = round(red * 32 / sec / 1e9, 1)
            if l2_read_peak:
                d["l2_read_frac_of_measured_gather_peak"] = round(rd * 32 / sec / 1e9 / l2_read_peak, 3)
            if l2_red_peak and red:
                d["l2_red_frac_of_measured_red_peak"] = round(red * 32 / sec / 1e9 / l2_red_peak, 3)
            d["dram_GBps"] = round(c["dram_bytes"] / sec / 1e9, 1)
            d["dram_frac_of_hbm_peak"] = round(c["dram_bytes"] / sec / 1e9 / hbm_peak, 3)
            d["dram_bytes_per_launch"] = c["dram_bytes"]
            d["tensor_pipe_pct_ncu"] = c.get("tensor_pipe_pct")
            d["issue_active_pct_ncu"] = c.get("issue_active_pct")
            d["top_stalls_ncu"] = c.get("top_stalls")
            d["counters_from"] = c.get("source")
            if c.get("note"):
                d["counters_note"] = c["note"]
        return d

    detail = {k: kernel_detail(k) for k in sorted(per_step, key=per_step.get, reverse=True)[:6]}
    top_d = detail[top]
    achieved = top_d.get("algorithmic_GBps")
    # what binds the dominant kernel according to its ncu capture (not according to the formula below)
    if top_d.get("dram_frac_of_hbm_peak", 0) > 0.6:
        bound = "hbm"
    elif (top_d.get("tensor_pipe_pct_ncu") or 0) > 50:
        bound = "tensor"
    else:
        bound = "latency"
    roofline = {
        "kernel": top, "bound": bound, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
        "frac": (achieved / hbm_peak) if achieved else None,
        "traffic": top_d.get("dram_bytes_per_launch"),
        "algorithmic_bytes_per_launch": alg_bytes.get(top),
        "peak_source": peak_src,
        "note": "achieved / frac follow the contract's recipe: ALGORITHMIC table bytes (SURVEY 8d: 512 B gathered or scattered per "
                "sample, each 2-feature RED counted as 4 B like the reference's half2 atomics) / live CUDA-event time / "
                "measured HBM copy peak. The table traffic is served by L1/L2, and the kernel executes fewer bytes than "
                "that (zero rows are skipped, cell runs merged): `by_kernel` gives the executed L2 sector rate against the "
                "MEASURED L2 gather / RED peaks (profiles/r3_l2_ubench.json), DRAM bytes against the HBM peak and the ncu "
                "tensor-pipe share. `bound` is what the ncu capture says: 'latency' = neither memory system nor tensor "
                "pipe above 60 % / 50 %; the kernel waits on tcgen05 round trips (top stall reasons listed).",
        "by_kernel": detail,
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
    losses.clear()
    ms_e2e = timed(K, e2e_step) / K
    check_losses("end-to-end run")
    staged.clear()
    e2e = {"value": world * B * 1e3 / ms_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
           "ms_per_step": ms_e2e}

    # ---- strong scaling (N > 1): the SAME global batch split over the ranks --------------------------
    # The weak-scaling `value` above keeps 2^18 rays per GPU, where the gradient all-reduce (191 MB fp32)
    # is ~1 % of the step. Here the global batch is fixed: `--rays` rays (and the shipped 8192 rays of
    # configs/instant_ngp.json) per step over all ranks, so the per-GPU compute shrinks with N while the
    # all-reduce and the dense AdamW do not. Efficiency is for the driver / reader to compute from the
    # N = 1 line's ms_per_step (the same global batch on one GPU is the N = 1 `value` itself).
    strong = None
    if world > 1:
        strong = {}
        for label, glob in (("global_rays_2^18", B), ("global_rays_8192", 8192)):
            per = glob // world
            if per < 1:
                continue
            shard = [{k: v[:per].contiguous() for k, v in b.items()} for b in batches]
            for i in range(3):
                step(shard[i % nb], shard[(i + 1) % nb])
            losses.clear()
            n_s = K if glob == B else 10 * K
            ms_s = timed(n_s, lambda i: step(shard[i % nb], shard[(i + 1) % nb])) / n_s
            check_losses(f"strong scaling {label}")
            strong[label] = {"global_rays": per * world, "rays_per_gpu": per, "ms_per_step": ms_s,
                             "value": per * world * 1e3 / ms_s, "unit": UNIT}
        strong["note"] = ("same global batch over all ranks; gradients all-reduced (sum) over NCCL and averaged in the "
                          "fused AdamW; compare ms_per_step with the N = 1 run at the same global batch")

    # ---- extraction: voxel queries/s (BASELINE.json metric, second half), inputs resident in HBM ----
    # SURVEY 8d workload: a dense grid of voxel columns over the granule x 81 altitudes (0..20 km, step
    # 250 m), float64 ECEF points normalised like scripts/extract.py:81, one call = 32768 columns x 81.
    # Each rank queries its own slab of columns (no communication).
    from atmonr.datasets.harp2_extract import HARP2VoxelGridExtractDataset
    n_vox = 32768 * 81
    lat_span_m = math.radians(float(dataset.lat[~dataset.lat.isnan()].max() - dataset.lat[~dataset.lat.isnan()].min())) * 6378137.0
    h_step = lat_span_m / 620.0   # ~620 x ~500 columns over the 5 x 5 degree granule: > 4 x 32768 columns
    grid_ds = HARP2VoxelGridExtractDataset(dataset, horizontal_step=h_step, alt_step=250.0, layout="regular")
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
    # SURVEY 8d per voxel: 512 B gathered + 24 B in (float64 xyz) + 4 B out; 1536 MAC on the tensor cores
    ext_c = counters.get("atmonr_extract_sigma_tc", {})
    ext_gbs = 540.0 * n_vox / (ms_ext * 1e-3) / 1e9
    extract = {"value": world * n_vox * 1e3 / ms_ext, "unit": "voxels/s", "voxels_per_call_per_gpu": n_vox,
               "ms_per_call": ms_ext,
               "workload": f"voxel grid, {per_call} columns x {n_alt} altitudes per call, float64 points (scripts/extract.py voxelgrid mode)",
               "roofline": {"kernel": "atmonr_extract_sigma_tc", "bound": "latency", "achieved": ext_gbs, "peak": hbm_peak,
                            "unit": "GB/s", "frac": ext_gbs / hbm_peak, "algorithmic_bytes_per_voxel": 540,
                            "traffic": ext_c.get("dram_bytes") if ext_c.get("workload_voxels") == n_vox else None,
                            "executed_l2_read_frac_of_measured_gather_peak":
                                (round(ext_c["l2_sectors_read"] * 32 / (ms_ext * 1e-3) / 1e9 / l2_read_peak, 3)
                                 if ext_c.get("l2_sectors_read") and l2_read_peak and ext_c.get("workload_voxels") == n_vox else None),
                            "tensor_pipe_pct_ncu": ext_c.get("tensor_pipe_pct"), "issue_active_pct_ncu": ext_c.get("issue_active_pct"),
                            "note": "algorithmic table bytes (L1/L2-served) against the HBM copy peak, as for the training kernels; "
                                    "the float64 geodesy in front (FP64 pipe) and the gathers' issue slots bound it"}}

    nerf = gpu_nerf_rate(dataset, dev)          # every rank (data-parallel NeRF line at N > 1)
    if rank != 0:
        return
    # the CPU baselines belong to the N = 1 run (rank 0 would otherwise keep the other ranks' GPUs idle
    # for ~20 s of host work in every run of the scaling sweep)
    solo = world == 1
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
            "gradient_exchange": ("hash tables: reduce-scatter -> AdamW on 1/world of the entries -> all-gather of the fp16 "
                                  "shadow; MLPs: all-reduce" if dp_sharded else ("all-reduce" if world > 1 else "none")),
            "backward": (f"samples with a non-zero incoming gradient only ({active_fraction:.3f} of all samples in the last step; exact)"
                         if _fused.COMPACT_BWD else "dense (every sample)"),
            "l2": "inputs larger than L2: per-step working set (x01, sigma, colour, gradients) is several GB",
        },
        "loss_first_last": [step_losses[0], step_losses[-1]],
        "clocks": clock_info, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "extract": extract,
        "strong_scaling": strong,
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
