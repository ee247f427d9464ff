"""Trainer host logic (reference: src/atmonr/trainer.py:26-274) on CPU with a toy pipeline: epochs,
the `fixed` and `target_lr` schedules, loss read-back in groups, progress pixels, checkpoint keys and
resume. The real pipelines need a GPU; everything the Trainer itself does is torch/numpy."""

import json
import os
from pathlib import Path

import pytest
import torch

from helpers import ROOT


class ToyPipeline:
    """One colour per band; renders it for every ray. Enough surface for the Trainer."""

    def __init__(self):
        self.color = torch.nn.Parameter(torch.tensor([0.05, 0.1, 0.15, 0.2]))
        self.device = "cpu"
        self.forward_calls = 0

    def get_optimizer(self, cfg):
        return torch.optim.SGD([self.color], lr=cfg["lr"])

    def forward(self, batch):
        self.forward_calls += 1
        cm = self.color[None].expand(batch["origin"].shape[0], 4)
        return {"color_map_fine": cm, "color_map_surf": 0.25 * cm, "color_map_atmo": 0.75 * cm}

    def compute_loss(self, batch, results):
        pred = torch.take_along_dim(results["color_map_fine"], batch["irgb_idx"][:, None], dim=1)[:, 0]
        return ((pred - batch["rad"]) ** 2).mean()

    def state_dict(self):
        return {"toy": {"color": self.color.detach().clone()}}

    def load_state_dict(self, sd):
        with torch.no_grad():
            self.color.copy_(sd["toy"]["color"])

    def train(self): ...
    def eval(self): ...


class Writer:
    def __init__(self):
        self.scalars, self.images = [], []

    def add_scalar(self, tag, val, step):
        self.scalars.append((tag, float(val), int(step)))

    def add_image(self, tag, img):
        self.images.append((tag, img.shape))


def _trainer(monkeypatch, tmp_path, sched, num_iters=14):
    from atmonr import trainer as T
    from atmonr.datasets.harp2 import HARP2Dataset
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(T, "_make_writer", lambda d: Writer())
    cfg = json.load(open(os.path.join(ROOT, "configs", "instant_ngp.json")))
    ds = HARP2Dataset(dict(cfg["dataset"]), "synthetic:H=12,W=12,seed=1", device=torch.device("cpu"))
    tcfg = dict(cfg["trainer"], batch_size=4096, num_iters=num_iters, print_frequency=4,
                optimizer={"lr": 0.5}, scheduler=sched)
    pipe = ToyPipeline()
    return T.Trainer(tcfg, ds, pipe, "toy"), pipe, ds


def test_fixed_schedule_checkpoints_and_resume(monkeypatch, tmp_path):
    sched = {"type": "fixed", "gamma": 0.5, "decay_start": 4, "decay_interval": 3}
    tr, pipe, ds = _trainer(monkeypatch, tmp_path, sched)
    per_epoch = len(tr.dataloader)
    assert per_epoch == -(-len(ds) // 4096) and tr.num_epochs == -(-14 // per_epoch)
    out = Path(tmp_path) / "run"
    out.mkdir()
    tr.train(out)
    assert tr.iter_count == 14 and pipe.forward_calls == 14
    # trainer.py:114-120: a decay every `decay_interval` iterations once iter_count > decay_start: 6, 9, 12
    assert tr.optimizer.param_groups[0]["lr"] == pytest.approx(0.5 * 0.5 ** 3)
    # every step's loss reaches the writer with its own step index, in order; the loss goes down
    losses = [(s, v) for tag, v, s in tr.writer.scalars if tag == "Loss"]
    assert [s for s, _ in losses] == list(range(14)) and losses[-1][1] < losses[0][1]
    # one image + metrics + checkpoint per (possibly partial) epoch
    n_epochs = -(-14 // per_epoch)
    assert tr.epoch_idx == n_epochs and len(tr.writer.images) == n_epochs
    assert {t for t, _, _ in tr.writer.scalars} >= {"Loss", "PSNR_mean", "SSIM_mean"}
    ckpts = sorted(out.glob("epoch_*.pt"))
    assert [c.name for c in ckpts] == [f"epoch_{k:04d}.pt" for k in range(1, n_epochs + 1)]
    ck = torch.load(ckpts[-1], weights_only=False)
    assert set(ck) == {"pipeline", "optimizer", "scheduler", "tensorboard_dir", "epoch_idx", "iter_count"}   # trainer.py:245-254
    assert ck["iter_count"] == 14 and ck["epoch_idx"] == n_epochs
    # resume: a fresh trainer picks up the newest checkpoint (numeric order, not lexical)
    tr2, pipe2, _ = _trainer(monkeypatch, tmp_path, sched, num_iters=20)
    tr2.load(out)
    assert tr2.iter_count == 14 and tr2.epoch_idx == n_epochs
    assert torch.equal(pipe2.color.detach(), pipe.color.detach())
    assert tr2.optimizer.param_groups[0]["lr"] == pytest.approx(0.5 * 0.5 ** 3)
    tr2.train(out)
    assert tr2.iter_count == 20
    assert tr2.optimizer.param_groups[0]["lr"] == pytest.approx(0.5 * 0.5 ** 5)      # + iterations 15, 18


def test_target_lr_schedule(monkeypatch, tmp_path):
    tr, pipe, ds = _trainer(monkeypatch, tmp_path, {"type": "target_lr", "final_lr": 0.005})
    out = Path(tmp_path) / "run"
    out.mkdir()
    tr.train(out)
    # trainer.py:55-59,181-182: gamma = (final / initial) ** (1 / num_epochs), one step per epoch
    assert tr.optimizer.param_groups[0]["lr"] == pytest.approx(0.005, rel=1e-6)
    with pytest.raises(NotImplementedError):
        _trainer(monkeypatch, tmp_path, {"type": "cosine"})


def test_progress_pixels_follow_the_predictions(monkeypatch, tmp_path):
    """trainer.py:123-140: after an epoch every ray's pixel holds the prediction of its band (total,
    surface, atmosphere), scattered through the batch's ray index."""
    tr, pipe, ds = _trainer(monkeypatch, tmp_path, {"type": "fixed", "gamma": 1.0, "decay_start": 0, "decay_interval": 1},
                            num_iters=100)
    tr.config["optimizer"]["lr"] = 0.0
    for g in tr.optimizer.param_groups:
        g["lr"] = 0.0                                     # frozen colours: the expected image is known
    seen = {}
    orig = ds.get_progress_tracker

    def spy():
        seen["p"] = orig()
        return seen["p"]

    monkeypatch.setattr(ds, "get_progress_tracker", spy)
    tr.config["num_iters"] = len(tr.dataloader)           # exactly one epoch
    out = Path(tmp_path) / "run"
    out.mkdir()
    tr.train(out)
    want = pipe.color.detach()[ds.ray_irgb_idx].numpy()
    p = seen["p"]
    assert abs(p.pred_pixels - want).max() <= 1e-7
    assert abs(p.pred_pixels_surf - 0.25 * want).max() <= 1e-7 and abs(p.pred_pixels_atmo - 0.75 * want).max() <= 1e-7


@pytest.mark.parametrize("layout", ["regular", "vincenty"])
def test_extract_dataset_dump(tmp_path, layout):
    """The extract grid's output file (harp2_extract.py:429-596): without netCDF4 an .npz with the
    reference's variable names; shapes follow (along, across, vertical[, bands])."""
    import numpy as np
    from atmonr.datasets.factory import get_extract_dataset
    from atmonr.datasets.harp2 import HARP2Dataset
    cfg = json.load(open(os.path.join(ROOT, "configs", "instant_ngp.json")))["dataset"]
    ds = HARP2Dataset(dict(cfg), "synthetic:H=10,W=9,seed=4", device=torch.device("cpu"))
    grid = get_extract_dataset("voxelgrid", ds, horizontal_step=40000.0, alt_step=2000.0, layout=layout,
                               coord_mode="voxelgrid", extract_filename="x.nc")          # extra CLI keys are ignored
    rows, cols, n_alt = grid.shp
    assert n_alt == 11 and len(grid) == rows * cols * n_alt
    sigma = torch.rand(len(grid), 1)
    grid.dump(tmp_path / "ext.nc", sigma)
    try:
        import netCDF4  # noqa: F401
        return                                                   # the netCDF branch is exercised where the module exists
    except ImportError:
        pass
    out = np.load(tmp_path / "ext.npz")
    assert set(out.files) == {"extinction_coefficient", "latitude", "longitude", "height", "altitude",
                              "x_wgs84", "y_wgs84", "z_wgs84"}
    assert out["extinction_coefficient"].shape == (rows, cols, n_alt, 1)
    assert np.array_equal(out["extinction_coefficient"][..., 0].ravel(), sigma[:, 0].numpy())
    assert out["latitude"].shape == out["longitude"].shape == out["height"].shape == (rows, cols)
    assert out["altitude"].shape == (n_alt,) and out["x_wgs84"].shape == (rows, cols, n_alt)
    with pytest.raises(NotImplementedError):
        get_extract_dataset("octree", ds)
