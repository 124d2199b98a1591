"""Drop-in for chainer_maskrcnn/functions/roi_align/roi_align_2d.py.

Same names, argument meaning and tuple conventions as the reference operator:

  ROIAlign2D(outh, outw, spatial_scale)                 roi_align_2d.py:15-20
    .check_type_forward(in_types)                        :22-32
    .forward_gpu(inputs)  -> (top_data,)                 :90-146
    .backward_gpu(inputs, gy) -> (bottom_diff, None)     :192-281
    .forward_cpu / .backward_cpu                         :39-88 / :148-190
  roi_align_2d(x, rois, outh, outw, spatial_scale)      :284-307

``x`` is (N,C,H,W) float32, ``rois`` is (R,5) float32 [batch_index, x_min, y_min,
x_max, y_max] (:296-297); the result is (R,C,outh,outw) float32.

Differences, all additive:
  * arrays are torch CUDA tensors (chainer/cupy are not required) or NumPy
    arrays.  NumPy arrays are staged to the GPU and back: ``forward_cpu`` /
    ``backward_cpu`` keep their names for call-site compatibility but there is
    NO CPU arithmetic in this package;
  * ``sampling_ratio`` (default 1 = the reference's single bin-centre sample,
    coordinates rounded exactly like the reference NumPy path; >1 or <=0 follow
    the caffe2 semantics of the reference's C++ port, caffe2_roi_align.cpp);
  * device results use channels-last memory (logical shape unchanged);
  * calling the object (``ROIAlign2D(...)(x, rois)``) or ``roi_align_2d`` on
    tensors that require grad records a torch.autograd node, the analogue of
    chainer's Variable graph.
"""
import numpy as np
import torch

from ... import _engine, _host, _lib


class InvalidType(TypeError):
    """Stands in for chainer.utils.type_check.InvalidType (roi_align_2d.py:22-32)."""


class ROIAlign2D(object):
    """RoI align over a set of 2d planes."""

    def __init__(self, outh, outw, spatial_scale, sampling_ratio=1, options=None):
        self.outh, self.outw = outh, outw
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio
        self.options = options            # rpool_options fields by name (experiments)
        self._bottom_data_shape = None
        self._plan = None
        self._plan_key = None

    # -- type checking ------------------------------------------------------
    def check_type_forward(self, in_types):
        """in_types: the two input arrays (or anything with dtype/ndim/shape)."""
        if len(in_types) != 2:
            raise InvalidType("expected 2 inputs (x, rois), got %d" % len(in_types))
        x, rois = in_types
        x_dtype = _np_dtype(x)
        r_dtype = _np_dtype(rois)
        if x_dtype != np.float32:
            raise InvalidType("x.dtype == float32 expected, got %s" % x_dtype)
        if len(x.shape) != 4:
            raise InvalidType("x.ndim == 4 expected, got %d" % len(x.shape))
        if r_dtype != np.float32:
            raise InvalidType("rois.dtype == float32 expected, got %s" % r_dtype)
        if len(rois.shape) != 2:
            raise InvalidType("rois.ndim == 2 expected, got %d" % len(rois.shape))
        if rois.shape[1] != 5:
            raise InvalidType("rois.shape[1] == 5 expected, got %d" % rois.shape[1])

    # -- device path --------------------------------------------------------
    def forward_gpu(self, inputs):
        self.check_type_forward(inputs)
        bottom_data, bottom_rois = inputs
        self._bottom_data_shape = tuple(bottom_data.shape)
        (top,), self._plan = _engine.forward(
            [bottom_data], bottom_rois, None, [self.spatial_scale],
            [(self.outh, self.outw)], sampling_ratio=self.sampling_ratio,
            roi_format=_lib.ROI_XY, options=self.options)
        self._plan_key = _rois_key(bottom_rois)
        return top,

    def backward_gpu(self, inputs, gy):
        # Like the reference (retain_inputs((1,)), :92) only the RoIs and the
        # remembered input shape are needed; inputs[0] may be None.
        bottom_rois = inputs[1]
        plan = self._plan
        # the forward plan (per-RoI tables, schedule) is reused only for the very same RoI
        # tensor, unmodified since: same storage, shape and torch version counter
        if plan is None or self._plan_key != _rois_key(bottom_rois):
            plan = self._make_plan(bottom_rois)
        (bottom_diff,) = _engine.backward(plan, [gy[0]])
        return bottom_diff, None

    def _make_plan(self, bottom_rois):
        if self._bottom_data_shape is None:
            raise RuntimeError("backward called before forward: input shape unknown")
        return _engine.make_plan([self._bottom_data_shape], bottom_rois, None,
                                 [self.spatial_scale], [(self.outh, self.outw)],
                                 sampling_ratio=self.sampling_ratio, roi_format=_lib.ROI_XY,
                                 options=self.options)

    # -- host-array entry points (computed on the GPU) -----------------------
    def forward_cpu(self, inputs):
        self.check_type_forward(inputs)
        x, rois = inputs
        top, = self.forward_gpu((_host.h2d(x), _host.h2d(rois)))
        return _host.d2h(top),

    def forward_cpu2(self, inputs):
        """The reference's C++-port entry (roi_align_2d.py:34-37 ->
        caffe2_roi_align.forward, caffe2_roi_align.cpp:231-243): caffe2 sampling
        semantics with sampling_ratio fixed to 1 (:240), whatever this object's own
        sampling_ratio is.  Host arrays in, host array out, computed on the GPU."""
        self.check_type_forward(inputs)
        x, rois = inputs
        (top,), _ = _engine.forward([_host.h2d(x)], _host.h2d(rois), None, [self.spatial_scale],
                                    [(self.outh, self.outw)], sampling_ratio=1,
                                    coord_mode=_lib.COORD_CAFFE2, roi_format=_lib.ROI_XY)
        return _host.d2h(top),

    def backward_cpu(self, inputs, gy):
        rois = _host.h2d(inputs[1])
        g = _host.h2d(gy[0])
        if self._bottom_data_shape is None and inputs[0] is not None:
            self._bottom_data_shape = tuple(inputs[0].shape)
        plan = self._make_plan(rois)
        (bottom_diff,) = _engine.backward(plan, [g])
        return _host.d2h(bottom_diff), None

    # -- chainer.Function-style dispatch --------------------------------------
    def forward(self, inputs):
        if _host.is_host_array(inputs[0]):
            return self.forward_cpu(inputs)
        return self.forward_gpu(inputs)

    def backward(self, inputs, gy):
        if _host.is_host_array(gy[0]):
            return self.backward_cpu(inputs, gy)
        return self.backward_gpu(inputs, gy)

    def __call__(self, x, rois):
        if _host.is_host_array(x):
            return self.forward_cpu((x, rois))[0]
        self.check_type_forward((x, rois))
        self._bottom_data_shape = tuple(x.shape)
        (y,) = _engine.apply([x], rois, None, spatial_scales=[self.spatial_scale],
                             out_sizes=[(self.outh, self.outw)],
                             sampling_ratio=self.sampling_ratio, roi_format=_lib.ROI_XY)
        return y


def _rois_key(rois):
    return (rois.data_ptr(), tuple(rois.shape), tuple(rois.stride()), rois._version, str(rois.device))


def _np_dtype(a):
    d = a.dtype
    if isinstance(d, torch.dtype):
        return np.dtype(str(d).replace("torch.", ""))
    return np.dtype(d)


def roi_align_2d(x, rois, outh, outw, spatial_scale, sampling_ratio=1):
    """Spatial Region of Interest (ROI) align function (roi_align_2d.py:284-307).

    Args:
        x: (n: batch, c: channel, h: height, w: width) float32.
        rois: (n: data size, 5) float32, each row (batch_index, x_min, y_min,
            x_max, y_max).
        outh (int): height of the pooled output.
        outw (int): width of the pooled output.
        spatial_scale (float): scale by which the RoI is resized.
        sampling_ratio (int): samples per bin side (extension; default 1).

    Returns:
        (R, c, outh, outw) float32.
    """
    return ROIAlign2D(outh, outw, spatial_scale, sampling_ratio)(x, rois)
