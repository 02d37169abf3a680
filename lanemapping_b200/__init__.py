"""lanemapping_b200 -- B200-native MLS point cloud -> BEV raster stage for LaneMapping.

Layout: ``csrc/`` CUDA kernels + C-ABI (include/lm_bev.h); ``bev`` torch-facing host API;
``spec`` the frozen forward spec; ``synth`` seeded synthetic clouds; ``sidecar``/``las``/
``convert_data`` the offline cropped_tiff writer in the style of the reference's
data/convert_data.py; ``datasets``/``pcencoder`` the DATASETS / PCENCODER plug-ins;
``strips`` strip-sharded multi-GPU rasterisation.
"""
from .spec import (ACC_NAMES, ACC_PLANES, CH_DENSITY, CH_MAX_I, CH_MAX_Z, CH_MEAN_I, CH_MEAN_Z, CH_MIN_Z,
                   CHANNEL_NAMES, TILE, BevSpec)

__all__ = ["BevSpec", "TILE", "CH_MAX_I", "CH_MEAN_I", "CH_MIN_Z", "CH_MAX_Z", "CH_MEAN_Z", "CH_DENSITY",
           "CHANNEL_NAMES", "ACC_NAMES", "ACC_PLANES"]
__version__ = "0.1.0"
