"""Pipeline plugin base class (reference: src/atmonr/pipelines/pipeline.py). A pipeline owns
the state of one neural rendering algorithm, its loss and its optimizer; the Trainer and the
extract script only talk to this interface."""

from __future__ import annotations

import warnings
from typing import Any, Mapping

import torch
from torch.optim import Optimizer


class Pipeline:
    def __init__(self, config: dict, dataset) -> None:
        """pipeline.py:17-60: capture the scene frame and build the point preprocessor."""
        self.ray_origin_height = dataset.config["ray_origin_height"]
        if config["point_preprocessor"] == "horizontal" and config["include_height"]:
            raise AssertionError("include_height cannot be combined with the 'horizontal' point preprocessor")
        l_x = config.get("encoder", {}).get("L_x")
        if not config["point_preprocessor"] and isinstance(l_x, list) and len(set(l_x)) > 1:
            warnings.warn(
                "Are you sure you want to use a variable encoding dimension for non-transformed coordinates?"
            )
        self.device = -1
        self.config = config
        self.scale = dataset.scale
        self.offset = dataset.offset
        self.point_preprocessor = (
            dataset.get_point_preprocessor(config["point_preprocessor"]) if config["point_preprocessor"] else None
        )

    def send_tensors_to(self, device: int) -> None:
        raise NotImplementedError

    def get_optimizer(self, config: dict) -> Optimizer:
        raise NotImplementedError

    def forward(self, ray_batch: Mapping[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        raise NotImplementedError

    def extract(self, pts: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def compute_loss(self, ray_batch: Mapping[str, torch.Tensor], results: dict[str, torch.Tensor]) -> torch.Tensor:
        raise NotImplementedError

    def state_dict(self) -> Mapping[str, Mapping[str, Any]]:
        raise NotImplementedError

    def load_state_dict(self, state_dict: dict) -> None:
        raise NotImplementedError

    def parameters(self):
        """Every trainable tensor of the pipeline (not part of the reference interface: the
        data-parallel launchers broadcast rank 0's initial values through it)."""
        for name in getattr(self, "module_names", []):
            yield from getattr(self, name).parameters()
        for net in getattr(self, "nerf", {}).values():
            yield from net.parameters()

    def train(self) -> None:
        raise NotImplementedError

    def eval(self) -> None:
        raise NotImplementedError
