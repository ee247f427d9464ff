"""Data-parallel host logic on CPU with the gloo backend, world_size 2: equal ray shards, the
SAME permutation on every rank, summed gradients scaled by 1/world == single-process gradients."""

import os
import socket

import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Rays:
    def __init__(self, n):
        self.ray_idx = torch.arange(n, dtype=torch.int32)
        self.x = torch.arange(n, dtype=torch.float32)

    def __getbatch__(self, idx):
        return {"x": self.x[idx], "idx": self.ray_idx[idx]}


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "atmospheric-neural-rendering_b200"))
    from atmonr import distributed as dist
    from atmonr.batch_loader import BatchLoader

    r, w, _ = dist.init_from_env("gloo")
    assert (r, w) == (rank, world) and dist.is_active()
    ds = _Rays(1000)
    loader = BatchLoader(ds, batch_size=128, shuffle=True, rank=r, world_size=w, seed=5)
    seen = [b["idx"].clone() for b in loader]
    # a toy "pipeline": loss = mean over the shard of (p * x)^2 ; gradients must match the global mean
    p = torch.nn.Parameter(torch.tensor([0.5, -1.5]))
    opt = torch.optim.SGD([p], lr=0.1)
    first = next(iter(BatchLoader(ds, batch_size=128, shuffle=True, rank=r, world_size=w, seed=5)))
    loss = ((p[0] * first["x"] + p[1]) ** 2).mean()
    loss.backward()
    dist.all_reduce_gradients(opt)
    # rank 0's initial parameters reach every rank, for both pipeline families (the NeRF nets are plain
    # torch modules, so this part of the launcher logic runs on CPU)
    sys.path.insert(0, os.path.join(root, "tests"))
    sys.path.insert(0, root)
    from helpers import FakeDataset, tiny_scene
    import json
    from atmonr.pipelines.factory import get_pipeline
    torch.manual_seed(100 + rank)                      # different initial values per rank on purpose
    cfg = json.load(open(os.path.join(root, "configs", "nerf.json")))["pipeline"]
    cfg["mlp_hidden_dim"] = 16
    pipe = get_pipeline(cfg, FakeDataset(tiny_scene(h=2, w=2, n_views=3)))
    before = torch.cat([q.detach().flatten() for q in pipe.parameters()]).clone()
    dist.broadcast_parameters(pipe.parameters())
    after = torch.cat([q.detach().flatten() for q in pipe.parameters()]).clone()
    # progress pixels: each rank writes its own shard's predictions; after the merge every rank holds all
    pix = torch.full((3, 1000), -1.0)
    touched = torch.zeros(1000, dtype=torch.bool)
    for b in seen[:3]:
        ray = b.long()
        pix[:, ray] = torch.stack([ray.float(), 2 * ray.float(), 3 * ray.float()])
        touched[ray] = True
    merged = dist.merge_disjoint_updates(pix, touched)
    # ragged epochs: n % (batch * world) != 0, incl. tails smaller than the world size
    ragged = {}
    for n in (1000, 1001, 129, 130, 131, 257):
        ld = BatchLoader(_Rays(n), batch_size=128, shuffle=True, rank=r, world_size=w, seed=1)
        ragged[n] = (len(ld), [b["idx"].numel() for b in ld])
    torch.save({"seen": seen, "grad": p.grad.clone(), "first": first["idx"].clone(), "nerf_before": before,
                "nerf_after": after, "merged": merged, "touched": touched, "ragged": ragged}, out.format(rank))
    td.destroy_process_group()


def test_ray_sharding_and_gradient_allreduce(tmp_path):
    world, port = 2, _free_port()
    out = str(tmp_path / "rank{}.pt")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = torch.load(out.format(0)), torch.load(out.format(1))
    # shards are disjoint, equal-sized, and together cover every ray exactly once per epoch
    all_idx = torch.cat([torch.cat(r0["seen"]), torch.cat(r1["seen"])])
    assert all_idx.numel() == 1000 and torch.equal(torch.sort(all_idx)[0], torch.arange(1000, dtype=torch.int32))
    assert r0["seen"][0].numel() == r1["seen"][0].numel() == 64
    # all-reduced (and averaged) gradient == gradient of the mean loss over the union of the shards
    assert torch.allclose(r0["grad"], r1["grad"])
    x = torch.cat([r0["first"], r1["first"]]).float()
    p = torch.nn.Parameter(torch.tensor([0.5, -1.5]))
    ((p[0] * x + p[1]) ** 2).mean().backward()
    assert torch.allclose(r0["grad"], p.grad, rtol=1e-5)
    # progress pixels of both ranks' shards are on every rank after the merge
    assert torch.equal(r0["merged"], r1["merged"])
    both = r0["touched"] | r1["touched"]
    assert int(both.sum()) == 3 * 128 and not bool((r0["touched"] & r1["touched"]).any())
    ray = torch.arange(1000.0)
    assert torch.equal(r0["merged"][:, both], torch.stack([ray, 2 * ray, 3 * ray])[:, both])
    assert bool((r0["merged"][:, ~both] == -1).all())
    # ragged tails: same number of batches and the same (non-zero) shard size on every rank
    assert r0["ragged"] == r1["ragged"]
    for n, (length, sizes) in r0["ragged"].items():
        assert length == len(sizes) and all(sz > 0 for sz in sizes), (n, length, sizes)
    assert r0["ragged"][129] == (1, [64]) and r0["ragged"][130] == (2, [64, 1]) and r0["ragged"][131] == (2, [64, 1])
    # parameter broadcast: the ranks started from different values and end with rank 0's
    assert r0["nerf_before"].numel() > 1000 and not torch.equal(r0["nerf_before"], r1["nerf_before"])
    assert torch.equal(r0["nerf_after"], r0["nerf_before"]) and torch.equal(r1["nerf_after"], r0["nerf_before"])


def test_shard_slice_partitions():
    from atmonr.distributed import shard_slice
    for n in (0, 1, 7, 8, 1000):
        for w in (1, 2, 3, 8):
            parts = [shard_slice(n, r, w) for r in range(w)]
            covered = [i for s in parts for i in range(s.start, s.stop)]
            assert covered == list(range(n))


def _torch_adamw_step(param, grad, exp_avg, exp_avg_sq, param_f16, lr, beta1, beta2, eps, weight_decay, step,
                      grad_scale=1.0, zero_grad=False):
    """Stand-in for the CUDA kernel on CPU tensors (test infrastructure: the kernel itself is checked against
    torch.optim.AdamW on the GPU): torch's single-tensor AdamW arithmetic, in place, shadow refreshed."""
    g = grad * grad_scale
    param.mul_(1 - lr * weight_decay)
    exp_avg.lerp_(g, 1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    denom = (exp_avg_sq.sqrt() / (1 - beta2 ** step) ** 0.5).add_(eps)
    param.addcdiv_(exp_avg, denom, value=-lr / (1 - beta1 ** step))
    if param_f16 is not None:
        param_f16.copy_(param.half())
    if zero_grad:
        grad.zero_()


def _sharded_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "atmospheric-neural-rendering_b200"))
    from atmonr import distributed as dist
    from atmonr.native import ops
    from atmonr.native.modules import shadow_of
    from atmonr.optim import FusedAdamW
    ops.adamw_step = _torch_adamw_step
    dist.init_from_env("gloo")
    torch.manual_seed(0)
    big = torch.nn.Parameter(torch.randn(4096))      # "hash table": sharded
    small = torch.nn.Parameter(torch.randn(40))      # "MLP": all-reduced
    opt = FusedAdamW([{"params": [big], "weight_decay": 0.0}, {"params": [small], "weight_decay": 0.01}],
                     lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    opt.shard_large_parameters(min_numel=1024)
    assert opt.is_sharded(big) and not opt.is_sharded(small)
    shadows = []
    for it in range(3):
        g = torch.Generator().manual_seed(10 * it + rank)        # every rank its own gradient
        big.grad, small.grad = torch.randn(4096, generator=g), torch.randn(40, generator=g)
        dist.all_reduce_gradients(opt)
        opt.step()
        shadows.append(shadow_of(big).clone())
    stale = big.detach().clone()
    opt.consolidate()
    torch.save({"big": big.detach().clone(), "small": small.detach().clone(), "shadow": shadows, "stale": stale,
                "m": opt.state[big]["exp_avg"].clone(), "v": opt.state[big]["exp_avg_sq"].clone(),
                "state_keys": sorted(opt.state_dict()["state"][0])}, out.format(rank))
    td.destroy_process_group()


def test_sharded_optimizer_equals_the_replicated_one(tmp_path):
    """FusedAdamW.shard_large_parameters: reduce-scatter -> AdamW on this rank's slice -> all-gather
    of the fp16 shadow, against one process that applies AdamW to the mean of the two ranks' gradients; the
    shadow is complete on every rank after every step, the float32 master and the moments after consolidate()."""
    world, port = 2, _free_port()
    out = str(tmp_path / "sh{}.pt")
    mp.spawn(_sharded_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = torch.load(out.format(0)), torch.load(out.format(1))
    torch.manual_seed(0)
    big, small = torch.randn(4096), torch.randn(40)
    mb, vb, ms, vs = torch.zeros(4096), torch.zeros(4096), torch.zeros(40), torch.zeros(40)
    want_shadow = []
    for it in range(3):
        gs = [torch.Generator().manual_seed(10 * it + r) for r in range(2)]
        gb, gsm = [], []
        for g in gs:
            gb.append(torch.randn(4096, generator=g)); gsm.append(torch.randn(40, generator=g))
        _torch_adamw_step(big, gb[0] + gb[1], mb, vb, None, 1e-2, 0.9, 0.99, 1e-15, 0.0, it + 1, grad_scale=0.5)
        _torch_adamw_step(small, gsm[0] + gsm[1], ms, vs, None, 1e-2, 0.9, 0.99, 1e-15, 0.01, it + 1, grad_scale=0.5)
        want_shadow.append(big.half())
    for r in (r0, r1):
        assert torch.allclose(r["big"], big, rtol=1e-6, atol=1e-7) and torch.allclose(r["small"], small, rtol=1e-6, atol=1e-7)
        assert torch.allclose(r["m"], mb, rtol=1e-6, atol=1e-8) and torch.allclose(r["v"], vb, rtol=1e-6, atol=1e-10)
        for got, want in zip(r["shadow"], want_shadow):
            assert torch.equal(got, want)
        assert r["state_keys"] == ["exp_avg", "exp_avg_sq", "step"]
    # before consolidate() each rank's master copy was current on its own half only
    assert torch.allclose(r0["stale"][:2048], big[:2048], rtol=1e-6, atol=1e-7) and not torch.allclose(r0["stale"][2048:], big[2048:])
    assert torch.allclose(r1["stale"][2048:], big[2048:], rtol=1e-6, atol=1e-7) and not torch.allclose(r1["stale"][:2048], big[:2048])
