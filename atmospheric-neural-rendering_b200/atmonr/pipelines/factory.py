"""Pipeline registry (reference: src/atmonr/pipelines/factory.py)."""

from __future__ import annotations

from atmonr.pipelines.instant_ngp import InstantNGPPipeline
from atmonr.pipelines.nerf import NeRFPipeline
from atmonr.pipelines.pipeline import Pipeline

_PIPELINES = {"NeRF": NeRFPipeline, "InstantNGP": InstantNGPPipeline}


def get_pipeline(config: dict, dataset) -> Pipeline:
    kind = config["type"]
    if kind not in _PIPELINES:
        raise NotImplementedError(f"Pipeline '{kind}' is unrecognized!")
    return _PIPELINES[kind](config, dataset)
