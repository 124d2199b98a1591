"""chainer_maskrcnn_b200 -- B200-native (sm_100a) FPN multi-level RoIAlign,
behind the operator API of katotetsuro/chainer-maskrcnn.

    functions.roi_align.roi_align_2d   ROIAlign2D, roi_align_2d
    functions.roi_align_2d_yx          _roi_align_2d_yx
    functions.fpn_roi_align            fpn_roi_align (fused head-level dispatch)
    model.rpn                          map_rois_to_fpn_levels
    model.head                         FPNRoIPooling, FPNRoIKeypointPooling

All arithmetic runs in csrc/ (librpool_b200.so, C ABI in include/rpool_b200.h).
There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._build import build as build_extension  # noqa: F401
from .functions import (ROIAlign2D, roi_align_2d, _roi_align_2d_yx,  # noqa: F401
                        fpn_roi_align, fpn_roi_align_host)
from .model.rpn import map_rois_to_fpn_levels  # noqa: F401
from .model.head import FPNRoIPooling, FPNRoIKeypointPooling  # noqa: F401

__version__ = "0.1.0"
