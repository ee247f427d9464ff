"""Query a trained pipeline for the extinction coefficient on a voxel grid.

Same command line as the reference's scripts/extract.py, all four coordinate modes (voxelgrid, l1c,
globalgrid, earthcare: atmonr/datasets/harp2_extract.py). Under torchrun the voxel columns are split contiguously over the
ranks with no communication; rank 0 gathers and writes the file.
"""

import argparse
import json
from pathlib import Path
from types import SimpleNamespace

import _bootstrap  # noqa: F401
import torch
from tqdm import tqdm

from atmonr import distributed as dist
from atmonr.batch_loader import BatchLoader
from atmonr.datasets.factory import BANDS, get_dataset, get_extract_dataset
from atmonr.geospatial.spherical import EARTH_RADIUS
from atmonr.pipelines.factory import get_pipeline


def _comma_separated(text: str) -> list[int]:
    return [int(t) for t in text.split(",")]


def parse_args() -> argparse.Namespace:
    ap = argparse.ArgumentParser()
    ap.add_argument("--exp-name", type=str, required=True)
    ap.add_argument("--coord-mode", type=str, required=True, help="l1c, voxelgrid, globalgrid or earthcare")
    ap.add_argument("--extract-filename", type=str, required=True)
    ap.add_argument("--batch-size", type=int, default=32768, help="voxel columns per batch")
    ap.add_argument("--min-alt", type=float)
    ap.add_argument("--max-alt", type=float)
    ap.add_argument("--alt-step", type=float, default=250.0)
    ap.add_argument("--horizontal-step", type=float, default=3000.0)
    ap.add_argument("--scale", type=float, default=100 / EARTH_RADIUS)
    ap.add_argument("--grid-res", type=float, default=0.025)
    ap.add_argument("--vstretch", type=float, default=12)
    ap.add_argument("--lon-crop", type=float, default=0.05)
    ap.add_argument("--earthcare-filename", type=str)
    ap.add_argument("--earthcare-range", type=_comma_separated)
    return ap.parse_args()


def main() -> None:
    args = parse_args()
    output_path = Path(f"data/output/{args.exp_name}")
    train_args = SimpleNamespace(**json.load(open(output_path / "args.json")))
    config = json.load(open(output_path / "config.json"))
    rank, world, local = dist.init_from_env()
    torch.cuda.set_device(local)
    device = torch.cuda.current_device()
    if not args.min_alt:
        args.min_alt = 0
    if not args.max_alt:
        args.max_alt = config["dataset"]["ray_origin_height"]

    dataset = get_dataset(config["dataset"], train_args.scene_filename)
    n_alt = torch.arange(args.min_alt, args.max_alt + args.alt_step / 2, args.alt_step).shape[0]
    extract_dataset = get_extract_dataset(args.coord_mode, dataset, **vars(args))
    pipeline = get_pipeline(config["pipeline"], dataset)
    pipeline.send_tensors_to(device)
    pipeline.eval()
    ckpts = sorted(output_path.glob("epoch_*.pt"), key=lambda c: int(c.stem.split("_")[1]))
    pipeline.load_state_dict(torch.load(ckpts[-1], weights_only=False)["pipeline"])

    num_bands = BANDS[config["dataset"]["type"]] if config["pipeline"].get("multi_band_extinction", False) else 1
    n_pts = extract_dataset.idx.shape[0]
    sigma = torch.zeros((n_pts, num_bands), device=device)
    # contiguous shard of voxel columns per rank (columns = groups of n_alt points; the earthcare and
    # globalgrid tables are not column-shaped and are sharded point by point)
    if args.coord_mode.lower() not in ("l1c", "voxelgrid"):
        n_alt = 1
    n_cols = n_pts // n_alt
    cols = dist.shard_slice(n_cols, rank, world)
    lo, hi = cols.start * n_alt, cols.stop * n_alt
    step = args.batch_size * n_alt
    with torch.no_grad():
        for start in tqdm(range(lo, hi, step), disable=rank != 0):
            sl = slice(start, min(start + step, hi))
            pts = (extract_dataset.xyz[sl] - dataset.offset) / dataset.scale
            sigma[sl] = pipeline.extract(pts).to(sigma.dtype) / dataset.scale
    if world > 1:
        torch.distributed.all_reduce(sigma)  # disjoint shards: the sum assembles the grid
    if rank == 0:
        extract_dataset.dump(output_path / args.extract_filename, sigma)


if __name__ == "__main__":
    main()
