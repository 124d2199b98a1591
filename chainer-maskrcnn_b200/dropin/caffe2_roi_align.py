"""caffe2_roi_align -- zero-edit integration point for the unmodified reference.

The reference's ``ROIAlign2D.forward_cpu`` / ``forward_cpu2`` first try
``import caffe2_roi_align`` and, when that works, return
``caffe2_roi_align.forward(bottom_data, bottom_rois, outh, outw, spatial_scale)``
(chainer_maskrcnn/functions/roi_align/roi_align_2d.py:34-46).  The module of that name
the reference ships is a pybind11 build of caffe2_operation/caffe2_roi_align.cpp:231-248:
c-contiguous float32 arrays in (force-cast), RoI rows [batch, x1, y1, x2, y2], caffe2
sampling semantics with sampling_ratio fixed to 1 (:240), a fresh zero-initialised
(R, C, out_h, out_w) float32 array out, ``std::runtime_error`` -> Python exception on a
malformed RoI array (:130-132).

This file has the same name and the same ``forward``; put its directory on ``sys.path``
(``PYTHONPATH=.../chainer-maskrcnn_b200/dropin``) and the reference's CPU entry points
run on the B200 through librpool_b200.so without touching a line of the reference.
There is no CPU arithmetic here: without the library or a device the call raises.
"""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

__doc_module__ = "roi align operation ported from caffe2"      # the pybind module's m.doc()


def forward(bottom_data, bottom_rois, out_h, out_w, spatial_scale):
    """caffe2_roi_align.cpp:231-243.  NumPy (N,C,H,W) / (R,5) in, NumPy (R,C,out_h,out_w) out."""
    from chainer_maskrcnn_b200 import _engine, _host, _lib
    x = np.ascontiguousarray(bottom_data, dtype=np.float32)       # py::array::forcecast | c_style
    rois = np.ascontiguousarray(bottom_rois, dtype=np.float32)
    if x.ndim != 4:
        raise RuntimeError("bottom_data must have 4 dimensions")
    if rois.ndim != 2 or rois.shape[1] != 5:
        raise RuntimeError("invalid roi shape")                   # caffe2_roi_align.cpp:130-132
    (top,), _ = _engine.forward([_host.h2d(x)], _host.h2d(rois), None, [float(spatial_scale)],
                                [(int(out_h), int(out_w))], sampling_ratio=1,
                                coord_mode=_lib.COORD_CAFFE2, roi_format=_lib.ROI_XY)
    return np.ascontiguousarray(_host.d2h(_engine.to_nchw_contiguous(top)))
