"""locate_b200: the LocAtE generator/discriminator training hot path on B200 (sm_100a).

Public surface mirrors /root/reference/libs/__init__.py:1-10 (same names, same call signatures) so the
reference's main.py can `import locate_b200 as libs`.  All arithmetic runs in the hand-written CUDA
kernels of locate_b200/csrc through the C ABI in include/locate_b200.h; there is no CPU fallback.
"""
from . import config, dist, ops
from ._lib import LIB_PATH, LocateLibraryError, launch_count, reset_launch_count
from .config import CFG, configure
from .layers import (ActivatedBaseConv, Block, BlockBlock, CatModule, DeepResidualConv, Expand, FeaturePooling,
                     InPlaceNorm, LinearModule, NonLinear, Norm, ResModule, RootTanhModule, Scale, SelfAttention,
                     SpectralNorm, feature_attention, nonlinear_function, residual_function)
from .models import Discriminator, Generator
from .optim import Nadam
from .train import GanTrainer, get_model, hinge, init, parameter_count, penalty

__all__ = [
    "NonLinear", "BlockBlock", "penalty", "Discriminator", "Generator", "SpectralNorm", "get_model", "hinge",
    "parameter_count", "Nadam", "GanTrainer", "CFG", "configure",
]
from .augment import GpuAugment  # noqa: E402,F401  (SURVEY "next" row N4: GPU-side input pipeline)
