"""atmonr -- B200-native implementation of the AtmoNR training / extraction hot path.

Same module paths and call surface as nasa/atmospheric-neural-rendering's `atmonr` package
(samplers, encoders, graphics_utils, losses, models.nerf, pipelines, datasets, trainer,
batch_loader, utils, geospatial) so `scripts/train.py` and `scripts/extract.py` run unchanged;
the per-step arithmetic runs in hand-written sm_100a kernels (libatmonr_b200.so, C ABI in
include/atmonr_b200.h) instead of eager PyTorch + tiny-cuda-nn.
"""

__version__ = "0.1.0"
