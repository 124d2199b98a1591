"""chainer_maskrcnn_b200 -- B200-native (sm_100a) FPN multi-level RoIAlign,
behind the operator API of katotetsuro/chainer-maskrcnn.

    functions.roi_align.roi_align_2d   ROIAlign2D, roi_align_2d
    functions.roi_align_2d_yx          _roi_align_2d_yx
    functions.fpn_roi_align            fpn_roi_align (fused head-level dispatch)
    model.rpn                          map_rois_to_fpn_levels
    model.head                         FPNRoIPooling, FPNRoIKeypointPooling
    FusedStep                          plan + forward + backward for static buffers, CUDA-graph replay
    chainer_adapter                    chainer.Function subclass over CuPy arrays (needs chainer)

All arithmetic runs in csrc/ (librpool_b200.so, C ABI in include/rpool_b200.h).
There is no CPU fallback.

The ctypes binding (``chainer_maskrcnn_b200._lib``) depends on nothing but the
standard library, so a CuPy/Chainer host can use it without torch: the
torch-typed adapters below are imported on first use.
"""
from . import _lib  # noqa: F401
from ._build import build as build_extension  # noqa: F401

__version__ = "0.2.0"

_LAZY = {
    "ROIAlign2D": ".functions", "roi_align_2d": ".functions", "_roi_align_2d_yx": ".functions",
    "fpn_roi_align": ".functions", "fpn_roi_align_host": ".functions",
    "map_rois_to_fpn_levels": ".model.rpn",
    "FPNRoIPooling": ".model.head", "FPNRoIKeypointPooling": ".model.head",
    "FusedStep": "._step",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        value = getattr(importlib.import_module(_LAZY[name], __name__), name)
        globals()[name] = value
        return value
    raise AttributeError("module %r has no attribute %r" % (__name__, name))


def __dir__():
    return sorted(list(globals()) + list(_LAZY))
