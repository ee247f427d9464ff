"""Train a neural rendering pipeline on a multi-angle satellite granule.

Same command line as the reference's scripts/train.py (--exp-name --config-path --scene-filename
--profile --overwrite --resume) and the same outputs under data/output/<exp-name>/. Besides
HARP2 netCDF granules, --scene-filename accepts `synthetic:H=..,W=..,seed=..` (built-in
generator). Under `torchrun --nproc-per-node N` the rays of every batch are sharded over N GPUs
and gradients are all-reduced over NCCL (the reference is single-GPU).
"""

import argparse
import json
import os
from pathlib import Path

import _bootstrap  # noqa: F401
import torch

from atmonr import distributed as dist
from atmonr.datasets.factory import get_dataset
from atmonr.pipelines.factory import get_pipeline
from atmonr.trainer import Trainer
from atmonr.utils import load_config


def parse_args() -> argparse.Namespace:
    ap = argparse.ArgumentParser()
    ap.add_argument("--exp-name", type=str, required=True, help="Name of this experiment.")
    ap.add_argument("--config-path", type=str, required=True, help="Path to the configuration for this experiment.")
    ap.add_argument("--scene-filename", type=str, required=True, help="Filename of the scene to reconstruct.")
    ap.add_argument("--profile", action="store_true", help="Use the pytorch profiler to analyze code performance.")
    ap.add_argument("--overwrite", action="store_true", help="Overwrite experiment directory if it exists.")
    ap.add_argument("--resume", action="store_true", help="Resume an interrupted experiment on the next epoch.")
    return ap.parse_args()


def setup_dir(args: argparse.Namespace, config: dict) -> Path:
    """The existence checks and the mkdir run on rank 0 only; the other ranks wait for its verdict
    (a slower rank must not trip over the directory rank 0 has just created)."""
    out = Path(f"data/output/{args.exp_name}")
    error = ""
    if dist.rank() == 0:
        if args.resume and not out.exists():
            error = f"--resume: {out} does not exist"
        elif not args.resume and not args.overwrite and out.exists():
            error = f"{out} exists (use --overwrite)"
        else:
            os.makedirs(out, exist_ok=True)
            with open(out / "args.json", "w") as fh:
                json.dump(vars(args), fh, indent=4)
            with open(out / "config.json", "w") as fh:
                json.dump(config, fh, indent=4)
    if dist.is_active():
        box = [error]
        torch.distributed.broadcast_object_list(box, src=0)   # also the barrier: the files exist afterwards
        error = box[0]
    assert not error, error
    return out


def main() -> None:
    args = parse_args()
    config = load_config(args.config_path)
    _, world, local = dist.init_from_env()
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    output_path = setup_dir(args, config)
    device = torch.cuda.current_device()
    dataset = get_dataset(config["dataset"], args.scene_filename)
    pipeline = get_pipeline(config["pipeline"], dataset)
    pipeline.send_tensors_to(device)
    if world > 1:
        dist.broadcast_parameters(pipeline.parameters())
    trainer = Trainer(config["trainer"], dataset, pipeline, args.exp_name)
    if args.resume:
        trainer.load(output_path)
    trainer.train(output_path, profile=args.profile)


if __name__ == "__main__":
    main()
