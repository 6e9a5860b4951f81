"""pifu_b200 - B200-native reconstruction hot path for RGB-D PIFuHD.

Drop-in for the reference's ``PIFuNetwNML`` / ``PIFuMRNet`` ``query``/``get_preds`` and
``mesh_util.reconstruction`` / ``eval_grid`` / ``eval_grid_octree``.  All
compute goes through ``libpifu_b200.so`` (hand-written sm_100a CUDA behind a C ABI,
``include/pifu_b200.h``); there is no CPU fallback - using a net without the built
library raises.
"""
__version__ = "0.1.0"

from . import config, synthetic                      # noqa: E402,F401
from .MLP import MLP                                 # noqa: E402,F401
from .PIFuNetwNML import PIFuNetwNML                 # noqa: E402,F401
from .PIFuMRNet import PIFuMRNet                     # noqa: E402,F401
from .engine import get_engine                       # noqa: E402,F401
