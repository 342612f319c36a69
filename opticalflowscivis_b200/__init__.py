"""opticalflowscivis_b200 — B200 (sm_100a) implementation of the OpticalFlowSciVis hot path.

    from opticalflowscivis_b200.flow3d.model.RIFE import Model          # Flow-3D/model/RIFE.py
    from opticalflowscivis_b200.flow2d.model.warplayer import warp      # Flow-2D/model/warplayer.py
    from opticalflowscivis_b200.upflow import CorrelationFunction       # UPFlow/model/correlation_package

All arithmetic runs in hand-written CUDA kernels behind the C ABI of `libofsv.so` (include/ofsv.h); PyTorch only
owns device memory and streams.  There is no CPU fallback.
"""
from . import ops                                         # noqa: F401
from .ops import reference_flavor, set_reference_flavor   # noqa: F401

__version__ = "0.1"
