#!/usr/bin/env bash
# usage (on the GPU box): bash scripts/validate_gpu.sh <tag>  -- tests, bench line, ncu launch list, ncu full captures
# (ATMONR_VALIDATE_QUICK=1: tests, bench line and CPU arm only -- the r4 pass of profiles/, ~6 GPU-minutes)
tag=${1:-r3}
o=gpurun_out
mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 $o/${tag}_pytest_gpu.log
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err || { tail -20 $o/${tag}_bench.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > $o/${tag}_bench_reference.json 2>> $o/${tag}_bench.err
[ -n "$ATMONR_VALIDATE_QUICK" ] && exit 0
export ATMONR_BENCH_NO_CLOCKS=1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_launch.log 2>&1
ATMONR_CUDA_PROFILER_RANGE=1 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_field|k_composite|k_ngp_sample|k_adamw' -c 8 -o $o/${tag}_prof_train \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_train.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_extract_sigma_tc' -s 3 -c 1 -o $o/${tag}_prof_extract \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_extract.log 2>&1
export ATMONR_NERF_STEPS=2
python scripts/bench_nerf.py > $o/${tag}_nerf_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/${tag}_nerf_launches.csv python scripts/bench_nerf.py > $o/${tag}_ncu_nerf_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_linear_tc|k_linear_dw' -s 170 -c 8 -o $o/${tag}_prof_nerf \
    python scripts/bench_nerf.py > $o/${tag}_ncu_nerf.log 2>&1
unset ATMONR_NERF_STEPS
ls -la $o; du -sh $o
python - <<PY
import json
d=json.loads(open("$o/${tag}_bench.json").read().strip().splitlines()[-1])
print("value %.3fM e2e %.3fM ms %.2f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"]), d["roofline"]["ms_per_step_by_kernel"], "extract %.2fG" % (d["extract"]["value"]/1e9), "nerf", d["nerf"]["value"], d["clocks"])
PY
