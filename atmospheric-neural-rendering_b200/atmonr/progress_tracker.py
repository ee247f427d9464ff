"""Host-side image buffers updated during training (reference: src/atmonr/progress_tracker.py)."""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import numpy.typing as npt


@dataclass
class ProgressTracker:
    valid: npt.NDArray[np.bool_]
    pred_img: npt.NDArray[np.float32]
    pred_img_surf: npt.NDArray[np.float32]
    pred_img_atmo: npt.NDArray[np.float32]
    pred_pixels: npt.NDArray[np.float32]
    pred_pixels_surf: npt.NDArray[np.float32]
    pred_pixels_atmo: npt.NDArray[np.float32]
    target_img: npt.NDArray[np.float32]
    target_img_rgb: npt.NDArray[np.float32]
