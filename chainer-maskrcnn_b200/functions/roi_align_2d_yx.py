"""Drop-in for chainer_maskrcnn/functions/roi_align_2d_yx.py:4-7.

The reference permutes the RoI columns [0,2,1,4,3] (ChainerCV's
(idx, y1, x1, y2, x2) -> the op's (idx, x1, y1, x2, y2)) with a fancy-index copy
before every call.  Here the kernel reads the yx order directly
(roi_format = RPOOL_ROI_YX): no copy, no extra launch.
"""
from .. import _engine, _host, _lib


def _roi_align_2d_yx(x, indices_and_rois, outh, outw, spatial_scale, sampling_ratio=1):
    if _host.is_host_array(x):
        (y,), _ = _engine.forward([_host.h2d(x)], _host.h2d(indices_and_rois), None,
                                  [spatial_scale], [(outh, outw)],
                                  sampling_ratio=sampling_ratio, roi_format=_lib.ROI_YX)
        return _host.d2h(y)
    (pool,) = _engine.apply([x], indices_and_rois, None, spatial_scales=[spatial_scale],
                            out_sizes=[(outh, outw)], sampling_ratio=sampling_ratio,
                            roi_format=_lib.ROI_YX)
    return pool
