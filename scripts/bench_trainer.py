"""Time the Trainer inner loop (forward, loss, backward, AdamW, progress bookkeeping) at the shipped
batch size of configs/instant_ngp.json (8192 rays x 1024 samples) on a synthetic granule:

    python scripts/bench_trainer.py [iterations]

Prints one JSON line: ms per iteration, rays/s, kernel launches per iteration. End-of-epoch work
(images, metrics, checkpoint) is excluded. DESIGN.md section 7 quotes its numbers."""
import sys, time, json, copy, torch
import os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, ROOT+'/atmospheric-neural-rendering_b200', ROOT+'/tests']
from pathlib import Path
from atmonr.datasets.factory import get_dataset
from atmonr.pipelines.factory import get_pipeline
from atmonr.trainer import Trainer
from atmonr.utils import load_config
from atmonr.native import lib as L
torch.cuda.set_device(0)
cfg=load_config(ROOT+'/configs/instant_ngp.json')
n_it=int(sys.argv[1]) if len(sys.argv)>1 else 300
cfg["trainer"]["num_iters"]=n_it
ds=get_dataset(cfg["dataset"], "synthetic:H=128,W=128,seed=0")
pipe=get_pipeline(cfg["pipeline"], ds); pipe.send_tensors_to(0)
tr=Trainer(cfg["trainer"], ds, pipe, "probe")
# monkeypatch end of epoch to nothing (we time the inner loop only)
tr._end_of_epoch=lambda *a, **k: None
import io, contextlib
torch.cuda.synchronize(); 
# warm
cfg["trainer"]["num_iters"]=30; tr.config["num_iters"]=30
with contextlib.redirect_stdout(io.StringIO()): tr.train(Path("/tmp/probe_out"))
torch.cuda.synchronize()
tr.config["num_iters"]=30+n_it
L.STATS=L.CallStats(timed=False)
t0=time.perf_counter()
with contextlib.redirect_stdout(io.StringIO()): tr.train(Path("/tmp/probe_out"))
torch.cuda.synchronize(); dt=time.perf_counter()-t0
print(json.dumps({"iters": n_it, "ms_per_iter": 1e3*dt/n_it, "rays_per_s": 8192*n_it/dt, "launches_per_iter": L.STATS.launches/n_it}))
