import sys, os, json, time
sys.path.insert(0, '.'); sys.path.insert(0, 'atmospheric-neural-rendering_b200'); sys.path.insert(0, 'tests')
import torch
import bench
from atmonr.datasets.factory import get_dataset
from atmonr.pipelines.factory import get_pipeline
from atmonr.batch_loader import BatchLoader
cfg = bench.pipeline_config(1024)
ds = get_dataset(cfg["dataset"], "synthetic:H=256,W=256,seed=0")
pipe = get_pipeline(cfg["pipeline"], ds); pipe.send_tensors_to(0)
opt = pipe.get_optimizer(bench.OPT_CFG)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
loader = BatchLoader(ds, batch_size=B, shuffle=True, seed=1)
batches = []
for b in loader:
    batches.append({k: b[k].contiguous() for k in ("origin", "dir", "len", "rad", "irgb_idx")})
    if len(batches) >= 4: break
def step(batch):
    res = pipe.forward(batch); loss = pipe.compute_loss(batch, res)
    opt.zero_grad(); loss.backward(); opt.step(); return loss
for i in range(3): step(batches[i % 4])
torch.cuda.synchronize()
def timed(K, sync_each=False):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(K):
        l = step(batches[i % 4])
        if sync_each: l.item()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K, (time.perf_counter() - t0) * 1e3 / K
print("nosync", timed(4)); print("nosync", timed(4)); print("sync", timed(4, True)); print("nosync", timed(4))
cs = bench.ClockSampler(0); cs.start(); print("nosync+smi", timed(4)); print(cs.stop())
print("mem", torch.cuda.max_memory_allocated() / 2**30, torch.cuda.memory_reserved() / 2**30)
print(torch.cuda.memory_stats()["num_alloc_retries"], torch.cuda.memory_stats()["num_device_alloc"], torch.cuda.memory_stats()["num_device_free"])
