"""Miscellaneous utilities (reference: src/atmonr/utils.py)."""

from __future__ import annotations

import json
from typing import Mapping

import torch


def load_config(config_path: str) -> dict:
    """utils.py:10-21: read the JSON config and normalise the casing of the two `type` keys."""
    with open(config_path) as fh:
        config = json.load(fh)
    canonical = {"pipeline": {"nerf": "NeRF", "instantngp": "InstantNGP"}, "dataset": {"harp2": "HARP2"}}
    for section, names in canonical.items():
        key = str(config[section]["type"]).lower()
        if key in names:
            config[section]["type"] = names[key]
    return config


def dict_to(d: dict[str, torch.Tensor], device) -> Mapping[str, torch.Tensor]:
    """utils.py:24-38: move every tensor of a dict to `device` (in place)."""
    for key, value in d.items():
        d[key] = value.to(device)
    return d
