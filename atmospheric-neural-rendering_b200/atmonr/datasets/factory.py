"""Dataset registry (reference: src/atmonr/datasets/factory.py)."""

from __future__ import annotations

from atmonr.datasets.harp2 import HARP2Dataset
from atmonr.datasets.harp2_extract import HARP2VoxelGridExtractDataset

BANDS = {"HARP2": 4}

Dataset = HARP2Dataset
ExtractDataset = HARP2VoxelGridExtractDataset

_DATASETS = {"HARP2": HARP2Dataset}
_EXTRACT_DATASETS = {"HARP2": {"voxelgrid": HARP2VoxelGridExtractDataset}}


def get_dataset(config: dict, filename: str) -> Dataset:
    kind = config["type"]
    if kind not in _DATASETS:
        raise NotImplementedError(f"Dataset '{kind}' is unrecognized!")
    return _DATASETS[kind](config, filename)


def get_extract_dataset(mode: str, dataset: Dataset, **kwargs) -> ExtractDataset:
    kind = dataset.config["type"]
    modes = _EXTRACT_DATASETS.get(kind, {})
    if mode not in modes:
        raise NotImplementedError(
            f"extract mode '{mode}' for dataset '{kind}' is outside this build's scope "
            f"(available: {sorted(modes)}; the L1C / EarthCARE / globalgrid modes are SURVEY 8f-4)"
        )
    return modes[mode](dataset, **kwargs)
