"""Turn the ncu artefacts a gpurun call brought back into the tracked summaries under profiles/.

    python scripts/summarise_ncu.py --tag r2 --full gpurun_out/X_prof_full.ncu-rep \\
        --launches gpurun_out/X_launches.csv --bench gpurun_out/X_bench.json --cmd-full "..." --cmd-launch "..."

Writes profiles/<tag>_ncu_full_summary.md (selected counters + warp-stall split per kernel),
profiles/<tag>_ncu_launch_list.md/.csv (per-kernel totals and shares next to the live CUDA-event
split of the bench line) and refreshes profiles/ncu_dram_traffic.json (read by bench.py for
roofline.traffic). Needs `ncu` on PATH (it only READS reports here; nothing is profiled).
"""

from __future__ import annotations

import argparse
import csv
import io
import json
import os
import subprocess
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__inst_executed_op_global_red.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
]
# kernel name prefix -> C-ABI entry point (the keys bench.py uses)
ENTRY = {
    "k_ngp_sample_points": "atmonr_ngp_sample_points", "k_field_fwd_tc": "atmonr_ngp_field_fwd_tc",
    "k_composite_fwd": "atmonr_composite_fwd", "k_composite_bwd": "atmonr_composite_bwd",
    "k_field_bwd_tc4": "atmonr_ngp_field_bwd_tc", "k_field_bwd_tc2": "atmonr_ngp_field_bwd_tc", "k_extract_sigma_tc": "atmonr_extract_sigma_tc",
    "k_adamw": "atmonr_adamw_step", "k_dense_tc": "atmonr_dense_fwd_tc", "k_dense_dw_tc": "atmonr_dense_dw_tc",
    "k_linear_tc": "atmonr_linear_fwd_tc", "k_linear_dw_tc": "atmonr_linear_dw_tc",
}


def short(name: str) -> str:
    name = name.replace("void ", "").replace("atm::", "")
    return name.split("(")[0]


def entry_of(kernel: str) -> str | None:
    for k, v in ENTRY.items():
        if kernel.startswith(k):
            # the last template argument of k_field_bwd_tc2<COMPACT> / k_composite_bwd<K, V, COMPACT>
            args = kernel[kernel.index("<") + 1:kernel.rindex(">")].split(",") if "<" in kernel else []
            compact = bool(args) and args[-1].strip() in ("1", "true") and (
                (k in ("k_field_bwd_tc2", "k_field_bwd_tc4")) or (k == "k_composite_bwd" and len(args) == 3))
            return v + ("_compact" if compact else "")
    return None


def full_summary(rep: str, tag: str, cmd: str, rays: int, write_traffic: bool = True) -> None:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    ki = head.index("Kernel Name")
    stall_cols = [i for i, c in enumerate(head) if c.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in c]
    out = [f"# ncu --set full, {tag}", "", f"`{cmd}`", "",
           "Durations under ncu are serialised, cache-flushed replays (longer than the live CUDA-event times of "
           "the bench line); counters are per launch. `stalls` = warp-state sampling split (pc sampling).", ""]
    traffic = {"rays": rays, "source": f"profiles/{tag}_ncu_full_summary.md (ncu --set full, dram__bytes_read.sum + "
                                        "dram__bytes_write.sum per launch)"}
    counters_path = os.path.join(ROOT, "profiles", "ncu_counters.json")
    counters = json.load(open(counters_path)) if os.path.exists(counters_path) else {}
    seen_ms: dict = {}     # entry -> longest ncu duration seen in THIS report (its counters are the ones kept)
    for r in data:
        name = short(r[ki])
        out += [f"## {name}", "```"]
        vals = {}
        for m in METRICS:
            if m in head:
                i = head.index(m)
                vals[m] = (r[i], units[i])
                out.append(f"{m:<78} {r[i]:>18} {units[i]}")
        tot = sum(float(r[i].replace(",", "") or 0) for i in stall_cols) or 1.0
        split = sorted(((float(r[i].replace(",", "") or 0) / tot, head[i][len("smsp__pcsamp_warps_issue_stalled_"):]) for i in stall_cols), reverse=True)
        out.append("stalls: " + ", ".join(f"{n} {100 * f:.1f}%" for f, n in split[:8]))
        out.append("```")
        out.append("")
        ent = entry_of(name)
        if ent and "dram__bytes_read.sum" in vals:
            def to_bytes(v, u):
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
                return float(v.replace(",", "")) * scale
            traffic[ent] = {"dram_bytes_per_launch": to_bytes(*vals["dram__bytes_read.sum"]) + to_bytes(*vals["dram__bytes_write.sum"])}

            def num(m):
                return float(vals[m][0].replace(",", "")) if m in vals and vals[m][0] not in ("", "n/a") else None
            dur_unit = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(vals["gpu__time_duration.sum"][1], 1e-6)
            this_ms = num("gpu__time_duration.sum") * dur_unit
            if this_ms < seen_ms.get(ent, 0.0):
                continue               # a shorter launch of the same entry point (smaller layer / tensor): keep the longest
            seen_ms[ent] = this_ms
            # per launch, from ONE ncu --set full capture; bench.py divides the byte / sector counts by the
            # LIVE CUDA-event duration of the same kernel (the ncu duration is a serialised cold-cache replay)
            counters[ent] = {
                "source": f"profiles/{tag}_ncu_full_summary.md", "workload_rays": rays,
                "workload_voxels": 32768 * 81 if ent.startswith("atmonr_extract") else None,
                "ncu_duration_ms": num("gpu__time_duration.sum") * dur_unit,
                "dram_bytes": traffic[ent]["dram_bytes_per_launch"],
                "l2_sectors_read": num("lts__t_sectors_srcunit_tex_op_read.sum"),
                "l2_sectors_red": num("lts__t_sectors_srcunit_tex_op_red.sum"),
                "l2_throughput_pct": num("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                "tensor_pipe_pct": num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
                "top_stalls": [f"{n} {100 * f:.1f}%" for f, n in split[:3]],
            }
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.md"), "w").write("\n".join(out))
    if write_traffic:
        json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_dram_traffic.json"), "w"), indent=1)
    json.dump(counters, open(counters_path, "w"), indent=1)


def launch_list(path: str, tag: str, cmd: str, bench: dict | None) -> None:
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum" or "atm::" not in r.get("Kernel Name", ""):
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        k = short(r["Kernel Name"])
        tot[k] += ms
        cnt[k] += 1
    total = sum(tot.values()) or 1.0
    live = (bench or {}).get("roofline", {}).get("ms_per_step_by_kernel", {})
    live_total = sum(live.values()) or 1.0
    out = [f"# ncu launch list, {tag} (steady state, own kernels only)", "", f"`{cmd}`", "",
           f"{sum(cnt.values())} launches of the library's kernels over the whole command (data-set set-up, settle and timed steps, "
           "extraction, the NeRF line); 2^18 rays x 1024 samples per Instant-NGP step. Times under ncu are serialised and cold-cache: compare SHARES with the live "
           "CUDA-event split of the same command without ncu (right-hand columns, the bench line of the same call).", "",
           "| kernel | launches | ncu total ms | ncu share | live ms/step | live share |", "|---|---|---|---|---|---|"]
    for k in sorted(tot, key=lambda k: -tot[k]):
        ent = entry_of(k)
        lv = live.get(ent) if ent else None
        if lv is None:
            lv = {"k_surface_bwd": live.get("atmonr_ngp_surface_bwd"), "k_surface_fwd": live.get("atmonr_ngp_surface_fwd"),
                  "k_adamw": live.get("atmonr_adamw_step"), "k_band_loss": live.get("atmonr_band_loss")}.get(k)
        out.append(f"| `{k}` | {cnt[k]} | {tot[k]:.3f} | {100 * tot[k] / total:.1f}% | "
                   + (f"{lv:.3f} | {100 * lv / live_total:.1f}% |" if lv is not None else " | |"))
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_launch_list.md"), "w").write("\n".join(out) + "\n")
    # the csv keeps the library's own kernels (namespace atm::); torch's set-up kernels are dropped
    own = [l for i, l in enumerate(lines) if i == 0 or "atm::" in l]
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_launch_list.csv"), "w").write("".join(own))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", required=True)
    ap.add_argument("--full")
    ap.add_argument("--launches")
    ap.add_argument("--bench")
    ap.add_argument("--cmd-full", default="")
    ap.add_argument("--cmd-launch", default="")
    ap.add_argument("--rays", type=int, default=1 << 18)
    ap.add_argument("--no-traffic", action="store_true", help="do not rewrite profiles/ncu_dram_traffic.json")
    a = ap.parse_args()
    bench = json.loads(open(a.bench).read().strip().splitlines()[-1]) if a.bench else None
    if a.full:
        full_summary(a.full, a.tag, a.cmd_full, a.rays, not a.no_traffic)
    if a.launches:
        launch_list(a.launches, a.tag, a.cmd_launch, bench)
    if bench:
        json.dump(bench, open(os.path.join(ROOT, "profiles", f"{a.tag}_bench_n1.json"), "w"))


if __name__ == "__main__":
    main()
