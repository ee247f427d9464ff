"""Training loop (reference: src/atmonr/trainer.py): same constructor, `train`, `save`, `load`,
checkpoint layout, schedulers and TensorBoard tags, so scripts/train.py runs unchanged.

Differences are host-side only: the three per-step progress-tracker copies are
gathered on the device and written to the host in one transfer, and the loss is read back once
per step instead of twice. Under torch.distributed (torchrun) gradients are all-reduced over
NCCL before the optimizer step (atmonr.distributed).
"""

from __future__ import annotations

import os

from datetime import datetime
from pathlib import Path

import numpy as np
import torch
from torch.optim.lr_scheduler import ExponentialLR
from torch.utils.data import DataLoader

from atmonr import distributed as dist
from atmonr.batch_loader import BatchLoader
from atmonr.utils import dict_to


class _NullWriter:
    def add_scalar(self, *a, **k): ...
    def add_image(self, *a, **k): ...


def _make_writer(log_dir):
    try:
        from torch.utils.tensorboard.writer import SummaryWriter  # noqa: PLC0415
        return SummaryWriter(log_dir)
    except Exception:  # tensorboard not installed
        return _NullWriter()


def _with_lookahead(loader):
    """(batch, next batch or None) pairs: the next batch is gathered one step early so that the
    pipeline can compute its sample points underneath the current step (Pipeline.prefetch)."""
    it = iter(loader)
    try:
        cur = next(it)
    except StopIteration:
        return
    for nxt in it:
        yield cur, nxt
        cur = nxt
    yield cur, None


class Trainer:
    def __init__(self, config: dict, dataset, pipeline, exp_name: str) -> None:
        self.config, self.dataset, self.pipeline = config, dataset, pipeline
        self.device = torch.cuda.current_device()
        self.rank, self.world_size = dist.rank(), dist.world_size()
        if hasattr(pipeline, "world_size"):
            pipeline.rank, pipeline.world_size = self.rank, self.world_size
        if config["all_gpu"]:
            assert config["num_workers"] == 0
            self.dataloader = BatchLoader(dataset, batch_size=config["batch_size"], shuffle=True,
                                          rank=self.rank, world_size=self.world_size)
        else:
            self.dataloader = DataLoader(dataset, num_workers=config["num_workers"],
                                         batch_size=config["batch_size"], shuffle=True)
        self.epoch_idx = 0
        self.iter_count = 0
        self.num_epochs = int(-(self.config["num_iters"] // -len(self.dataloader)))
        self.optimizer = pipeline.get_optimizer(self.config["optimizer"])
        if self.world_size > 1 and hasattr(self.optimizer, "shard_large_parameters") and os.environ.get("ATMONR_DP_SHARD", "1") != "0":
            # hash tables: reduce-scatter -> AdamW on 1/world of the entries -> all-gather of the fp16 shadow
            self.optimizer.shard_large_parameters()
        sched = self.config["scheduler"]
        if sched["type"] == "target_lr":
            gamma = (sched["final_lr"] / self.config["optimizer"]["lr"]) ** (1 / self.num_epochs)
        elif sched["type"] == "fixed":
            gamma = sched["gamma"]
        else:
            raise NotImplementedError(f"Unknown scheduler type {sched['type']}")
        self.scheduler = ExponentialLR(optimizer=self.optimizer, gamma=gamma)
        stamp = datetime.now().strftime("%Y%m%d_%H%M%S")
        self.tensorboard_dir = Path("data") / "tensorboard" / f"{exp_name}_{stamp}"
        self.writer = _make_writer(self.tensorboard_dir) if self.rank == 0 else _NullWriter()

    def train_step(self, batch):
        """forward -> loss -> zero_grad -> backward -> (all-reduce) -> step. trainer.py:99-105."""
        results = self.pipeline.forward(batch)
        loss = self.pipeline.compute_loss(batch, results)
        self.optimizer.zero_grad()
        loss.backward()
        dist.all_reduce_gradients(self.optimizer)
        self.optimizer.step()
        return results, loss

    def train(self, output_path: Path, profile: bool = False) -> None:
        prof = self.get_profiler() if profile else None
        if prof:
            prof.start()
        progress = self.dataset.get_progress_tracker()
        last_len, running = 0, []
        sched = self.config["scheduler"]
        # The reference reads the loss and three pixel vectors back to the host EVERY step
        # (trainer.py:108-140: two .item() and three device->host copies), which serialises host and
        # device. Same observable behaviour without the per-step synchronisation: losses
        # are read back in groups of `print_frequency` steps (the writer gets every step's value with
        # its own step index), and the progress pixels are scattered into a device buffer that is
        # copied into the host-side ProgressTracker once per epoch, before it is used.
        dev_pix = None           # (3, n_rays) on the device: total / surface / atmosphere predictions
        touched = None           # (n_rays,) rays this rank predicted in the current epoch (data-parallel merge)
        waiting: list = []       # (iteration, loss tensor) not yet read back

        def flush_losses():
            nonlocal running
            if not waiting:
                return
            vals = torch.stack([t for _, t in waiting]).float().cpu().tolist()
            for (it, _), v in zip(waiting, vals):
                self.writer.add_scalar("Loss", v, it)
            running = (running + vals)[-self.config["print_frequency"]:]
            waiting.clear()

        while self.iter_count < self.config["num_iters"]:
            for batch, upcoming in _with_lookahead(self.dataloader):
                if prof:
                    prof.step()
                if not self.config["all_gpu"]:
                    batch = dict_to(batch, self.device)
                elif upcoming is not None and hasattr(self.pipeline, "prefetch"):
                    self.pipeline.prefetch(upcoming)
                results, loss = self.train_step(batch)
                waiting.append((self.iter_count, loss.detach()))
                self.iter_count += 1
                if (sched["type"] == "fixed" and self.iter_count % sched["decay_interval"] == 0
                        and self.iter_count > sched["decay_start"]):
                    self.scheduler.step()
                # progress tracker (trainer.py:123-140): one gather + one scatter on the device
                band = batch["irgb_idx"][:, None]
                maps = torch.stack([results["color_map_fine"], results["color_map_surf"], results["color_map_atmo"]])
                pix = torch.take_along_dim(maps.detach(), band[None].expand(3, -1, -1), dim=2)[..., 0].float()
                if dev_pix is None:
                    dev_pix = torch.zeros((3, progress.pred_pixels.shape[0]), device=pix.device)
                    for k, name in enumerate(("pred_pixels", "pred_pixels_surf", "pred_pixels_atmo")):
                        dev_pix[k] = torch.from_numpy(getattr(progress, name)).to(pix.device)
                ray = batch["idx"].to(pix.device, torch.long)
                dev_pix.index_copy_(1, ray, pix)
                if self.world_size > 1:
                    if touched is None:
                        touched = torch.zeros(dev_pix.shape[1], device=pix.device, dtype=torch.bool)
                    touched[ray] = True
                if self.iter_count >= self.config["num_iters"]:
                    break
                if self.iter_count % self.config["print_frequency"] == 0:
                    flush_losses()
                    if self.rank == 0:
                        line = f"{self.iter_count}/{self.config['num_iters']} | Loss: {sum(running) / len(running):.5f}"
                        print(line + max(0, last_len - len(line)) * " ", end="\r")
                        last_len = len(line)
            flush_losses()
            if dev_pix is not None:
                if touched is not None:
                    # every rank predicted only its own ray shards: merge them so that rank 0's image,
                    # PSNR / SSIM and TensorBoard panels are those of the single-GPU run
                    dev_pix = dist.merge_disjoint_updates(dev_pix, touched)
                    touched.zero_()
                host_pix = dev_pix.cpu().numpy()
                progress.pred_pixels[:], progress.pred_pixels_surf[:], progress.pred_pixels_atmo[:] = host_pix
            self._end_of_epoch(progress, output_path, last_len)
            if prof:
                prof.stop()
                prof = None
        print()

    def _end_of_epoch(self, progress, output_path, last_len) -> None:
        """trainer.py:160-214: images, metrics, scheduler (target_lr), checkpoint."""
        dev = self.pipeline.device

        def as_cube(img, pixels):
            img[progress.valid] = pixels
            return torch.from_numpy(img).to(dev).permute(2, 0, 1)

        pred = as_cube(progress.pred_img, progress.pred_pixels)
        pred_surf = as_cube(progress.pred_img_surf, progress.pred_pixels_surf)
        pred_atmo = as_cube(progress.pred_img_atmo, progress.pred_pixels_atmo)
        target = torch.from_numpy(progress.target_img).to(dev)
        target[target.isnan()] = 0
        target = target.permute(2, 0, 1)
        self.epoch_idx += 1
        if self.config["scheduler"]["type"] == "target_lr":
            self.scheduler.step()
        if hasattr(self.optimizer, "consolidate"):
            self.optimizer.consolidate()     # collective: float32 masters / moments of all slices, before the checkpoint
        if self.rank != 0:
            return
        metrics = self.dataset.get_image_metrics(pred, target)
        line = f"Epoch {self.epoch_idx}/{self.num_epochs}"
        for name, val in metrics.items():
            if isinstance(val, list):
                continue
            line += f" | {name}: {val:.3f}"
            self.writer.add_scalar(name, val, self.epoch_idx)
        print(line + max(0, last_len - len(line)) * " ")
        rgb = lambda cube: self.dataset.get_rgb(cube).cpu().numpy()
        viz = np.concatenate([rgb(pred_surf), rgb(pred_atmo), rgb(pred), progress.target_img_rgb], axis=1)
        self.writer.add_image(f"Epoch {self.epoch_idx}", np.transpose(viz, (2, 0, 1)))
        self.save(output_path, self.epoch_idx)

    def get_profiler(self) -> torch.profiler.profile:
        return torch.profiler.profile(
            activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA],
            schedule=torch.profiler.schedule(wait=2, warmup=2, active=10, repeat=1),
            on_trace_ready=torch.profiler.tensorboard_trace_handler(str(self.tensorboard_dir)),
            record_shapes=True, profile_memory=True, with_stack=True, with_flops=True, with_modules=True,
        )

    def save(self, output_path: Path, epoch: int) -> None:
        """trainer.py:239-256: epoch_NNNN.pt with the same keys as the reference."""
        torch.save(
            {"pipeline": self.pipeline.state_dict(), "optimizer": self.optimizer.state_dict(),
             "scheduler": self.scheduler.state_dict(), "tensorboard_dir": self.tensorboard_dir,
             "epoch_idx": self.epoch_idx, "iter_count": self.iter_count},
            Path(output_path) / f"epoch_{epoch:04d}.pt",
        )

    def load(self, output_path: Path) -> None:
        """trainer.py:258-274: resume from the newest epoch_*.pt."""
        ckpts = sorted(Path(output_path).glob("epoch_*.pt"), key=lambda c: int(c.stem.split("_")[1]))
        ckpt = torch.load(ckpts[-1], weights_only=False)
        self.pipeline.load_state_dict(ckpt["pipeline"])
        self.optimizer.load_state_dict(ckpt["optimizer"])
        self.scheduler.load_state_dict(ckpt["scheduler"])
        self.tensorboard_dir = ckpt["tensorboard_dir"]
        self.writer = _make_writer(self.tensorboard_dir) if self.rank == 0 else _NullWriter()
        self.epoch_idx, self.iter_count = ckpt["epoch_idx"], ckpt["iter_count"]
        if hasattr(self.pipeline, "set_draw_counter"):
            self.pipeline.set_draw_counter(self.iter_count)
