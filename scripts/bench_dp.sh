#!/usr/bin/env bash
o=gpurun_out; mkdir -p $o
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $o/bench_dp$N.json 2> $o/bench_dp$N.err || { tail -30 $o/bench_dp$N.err; exit 1; }
ATMONR_DP_SHARD=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > $o/bench_dp${N}_allreduce.json 2> $o/bench_dp${N}_allreduce.err || { tail -30 $o/bench_dp${N}_allreduce.err; exit 1; }
python - <<PY
import json
for f in ("$o/bench_dp$N.json", "$o/bench_dp${N}_allreduce.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.3fM e2e %.3fM ms %.2f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"]), d["config"]["gradient_exchange"][:40], "strong", {k:(round(v["ms_per_step"],3), round(v["value"]/1e6,3)) for k,v in d["strong_scaling"].items() if isinstance(v,dict)}, "extract %.2fG" % (d["extract"]["value"]/1e9), "nerf", round(d["nerf"]["value"]), d["loss_first_last"])
PY
