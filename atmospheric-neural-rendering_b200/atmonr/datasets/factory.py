"""Dataset registry (reference: src/atmonr/datasets/factory.py)."""

from __future__ import annotations

from atmonr.datasets.harp2 import HARP2Dataset
from atmonr.datasets.harp2_extract import (HARP2EarthCAREExtractDataset, HARP2GlobalGridExtractDataset,
                                           HARP2L1CExtractDataset, HARP2VoxelGridExtractDataset)

BANDS = {"HARP2": 4}

Dataset = HARP2Dataset
ExtractDataset = HARP2VoxelGridExtractDataset

_DATASETS = {"HARP2": HARP2Dataset}
# datasets/factory.py:24-33 of the reference: the four coordinate modes of scripts/extract.py
_EXTRACT_DATASETS = {"HARP2": {"l1c": HARP2L1CExtractDataset, "voxelgrid": HARP2VoxelGridExtractDataset,
                               "globalgrid": HARP2GlobalGridExtractDataset, "earthcare": HARP2EarthCAREExtractDataset}}


def get_dataset(config: dict, filename: str) -> Dataset:
    kind = config["type"]
    if kind not in _DATASETS:
        raise NotImplementedError(f"Dataset '{kind}' is unrecognized!")
    return _DATASETS[kind](config, filename)


def get_extract_dataset(mode: str, dataset: Dataset, **kwargs) -> ExtractDataset:
    kind = dataset.config["type"]
    if kind not in _EXTRACT_DATASETS:
        raise NotImplementedError(f"ExtractDataset data_type '{kind}' is unrecognized!")
    modes = _EXTRACT_DATASETS[kind]
    if mode.lower() not in modes:
        raise NotImplementedError(f"extract mode '{mode}' for dataset '{kind}' is unrecognized (available: {sorted(modes)})")
    return modes[mode.lower()](dataset, **kwargs)
