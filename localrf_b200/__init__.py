"""localrf_b200 -- B200-native (sm_100a) render path of facebookresearch/localrf.

Public surface = the reference's own module API for this path:

    from localrf_b200 import LocalTensorfs, TensorVMSplit, AlphaGridMask

backed by the C ABI in include/localrf_b200.h (localrf_b200/csrc/liblrf_b200.so).  See DESIGN.md.
"""
from .tensorf import AlphaGridMask, MLPRender_Fea_late_view, TensorBase, TensorVMSplit  # noqa: F401
from .local_tensorfs import LocalTensorfs, ids2pixel, ids2pixel_view  # noqa: F401
from . import ray_utils, utils  # noqa: F401
from .pipeline import FramePipeline  # noqa: F401
from ._lib import build, lib  # noqa: F401

__version__ = "0.1.0"
