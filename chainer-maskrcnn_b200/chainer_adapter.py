"""Chainer-typed drop-in: the reference's operator class over librpool_b200.so.

``ROIAlign2D`` here has the reference's base class and conventions
(chainer_maskrcnn/functions/roi_align/roi_align_2d.py:15-20): an old-style
``chainer.function.Function`` whose ``forward_gpu(inputs)`` takes CuPy arrays
``(x, rois)``, calls ``self.retain_inputs((1,))`` (:92), remembers
``self._bottom_data_shape`` (:94) and returns the 1-tuple ``(top_data,)`` (:146);
``backward_gpu(inputs, gy)`` gets ``inputs[0] is None`` and returns
``(bottom_diff, None)`` (:192-195,281).  ``check_type_forward`` uses
``chainer.utils.type_check`` like the reference (:22-32).  ``roi_align_2d`` and
``_roi_align_2d_yx`` are the functional forms (:284-307, roi_align_2d_yx.py:4-7),
``fpn_roi_align`` replaces the heads' per-RoI loops (fpn_roi_mask_head.py:57-63).

Arrays stay CuPy's: outputs come from ``cupy.empty`` (:98-99), addresses are taken
with ``.data.ptr`` and launches go to ``cupy.cuda.get_current_stream()`` -- only the
ctypes binding ``_lib`` is used, torch is not imported.  The reference's arrays are
NCHW; they are converted to the kernels' channels-last layout by the library's own
transpose kernels (2 x the tensor per conversion).

Import-guarded: chainer and cupy are not installed in the image this repository is
developed in, so importing this module without them raises ImportError with that
message; tests/ exercise it against minimal stand-ins for both (tests/stub_chainer.py).
"""
import ctypes

try:
    import chainer
    from chainer import function
    from chainer.utils import type_check
    import cupy
except ImportError as _e:  # pragma: no cover - exercised through the stand-ins
    raise ImportError("chainer_maskrcnn_b200.chainer_adapter needs chainer and cupy (%s); the torch-typed "
                      "adapters in chainer_maskrcnn_b200.functions need neither" % _e)

import numpy

from . import _lib


def _stream():
    return ctypes.c_void_p(cupy.cuda.get_current_stream().ptr)


def _nhwc(a):
    """(N,C,H,W) CuPy array -> new (N,H,W,C) CuPy array (rpool_nchw_to_nhwc)."""
    a = cupy.ascontiguousarray(a)
    n, c, h, w = a.shape
    out = cupy.empty((n, h, w, c), dtype=numpy.float32)
    if a.size:
        _lib.check(_lib.lib().rpool_nchw_to_nhwc(a.data.ptr, out.data.ptr, n, c, h, w, _stream()))
    return out


def _nchw(a):
    """(N,H,W,C) CuPy array -> new (N,C,H,W) CuPy array (rpool_nhwc_to_nchw)."""
    n, h, w, c = a.shape
    out = cupy.empty((n, c, h, w), dtype=numpy.float32)
    if a.size:
        _lib.check(_lib.lib().rpool_nhwc_to_nchw(a.data.ptr, out.data.ptr, n, c, h, w, _stream()))
    return out


def _problem(level_arrays, scales, channels, rois, roi_format, levels, out_sizes, pooled, sampling_ratio,
             coord_mode=None):
    p = _lib.Problem()
    p.n_levels, p.channels = len(level_arrays), channels
    p.feat_layout = p.pool_layout = _lib.NHWC
    for l, (a, sc) in enumerate(zip(level_arrays, scales)):
        p.level[l].data = a.data.ptr
        p.level[l].n_images, p.level[l].height, p.level[l].width = a.shape[0], a.shape[1], a.shape[2]
        p.level[l].spatial_scale = float(sc)
    p.rois, p.n_rois, p.roi_format = rois.data.ptr, rois.shape[0], roi_format
    if levels is not None:
        if levels.dtype == numpy.int32:
            p.roi_levels = levels.data.ptr
        else:
            p.roi_levels_f32 = levels.data.ptr          # map_rois_to_fpn_levels returns float32
    p.n_heads = len(out_sizes)
    for h, ((oh, ow), o) in enumerate(zip(out_sizes, pooled)):
        p.out_h[h], p.out_w[h], p.pooled[h] = oh, ow, o.data.ptr
    p.sampling_ratio = sampling_ratio
    p.coord_mode = coord_mode if coord_mode is not None else \
        (_lib.COORD_CHAINER if sampling_ratio == 1 else _lib.COORD_CAFFE2)
    return p


def _workspace(p):
    n = _lib.lib().rpool_workspace_bytes_ex(p.n_rois, p.n_heads, p.coord_mode)
    return cupy.empty((n,), dtype=numpy.uint8)


class ROIAlign2D(function.Function):
    """RoI align over a set of 2d planes."""

    def __init__(self, outh, outw, spatial_scale, sampling_ratio=1):
        self.outh, self.outw = outh, outw
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio

    def check_type_forward(self, in_types):
        type_check.expect(in_types.size() == 2)
        x_type, roi_type = in_types
        type_check.expect(
            x_type.dtype == numpy.float32,
            x_type.ndim == 4,
            roi_type.dtype == numpy.float32,
            roi_type.ndim == 2,
            roi_type.shape[1] == 5,
        )

    def forward_gpu(self, inputs):
        self.retain_inputs((1,))
        self._bottom_data_shape = inputs[0].shape
        bottom_data, bottom_rois = inputs
        n_rois, channels = bottom_rois.shape[0], bottom_data.shape[1]
        x = _nhwc(bottom_data)
        rois = cupy.ascontiguousarray(bottom_rois)
        top = cupy.empty((n_rois, self.outh, self.outw, channels), dtype=numpy.float32)
        p = _problem([x], [self.spatial_scale], channels, rois, _lib.ROI_XY, None,
                     [(self.outh, self.outw)], [top], self.sampling_ratio)
        ws = _workspace(p)
        L = _lib.lib()
        _lib.check(L.rpool_plan(ctypes.byref(p), ws.data.ptr, ws.size, _stream()))
        _lib.check(L.rpool_forward(ctypes.byref(p), ws.data.ptr, ws.size, _stream()))
        return _nchw(top),

    def backward_gpu(self, inputs, gy):
        bottom_rois = inputs[1]                      # inputs[0] is None: retain_inputs((1,))
        n, channels, height, width = self._bottom_data_shape
        rois = cupy.ascontiguousarray(bottom_rois)
        g = _nhwc(gy[0])                             # (R,C,oh,ow) -> (R,oh,ow,C)
        diff = cupy.empty((n, height, width, channels), dtype=numpy.float32)
        p = _problem([diff], [self.spatial_scale], channels, rois, _lib.ROI_XY, None,
                     [(self.outh, self.outw)], [g], self.sampling_ratio)
        ws = _workspace(p)
        L = _lib.lib()
        _lib.check(L.rpool_plan(ctypes.byref(p), ws.data.ptr, ws.size, _stream()))
        _lib.check(L.rpool_backward(ctypes.byref(p), ws.data.ptr, ws.size, _stream()))     # zero fill inside
        return _nchw(diff), None

    # host arrays: staged through the GPU (there is no CPU arithmetic in this package)
    def forward_cpu(self, inputs):
        self._bottom_data_shape = inputs[0].shape
        top, = self.forward_gpu((cupy.asarray(inputs[0]), cupy.asarray(inputs[1])))
        return cupy.asnumpy(top),

    def backward_cpu(self, inputs, gy):
        if inputs[0] is not None:
            self._bottom_data_shape = inputs[0].shape
        diff, _ = self.backward_gpu((None, cupy.asarray(inputs[1])), (cupy.asarray(gy[0]),))
        return cupy.asnumpy(diff), None


def roi_align_2d(x, rois, outh, outw, spatial_scale, sampling_ratio=1):
    """roi_align_2d.py:284-307."""
    return ROIAlign2D(outh, outw, spatial_scale, sampling_ratio)(x, rois)


def _roi_align_2d_yx(x, indices_and_rois, outh, outw, spatial_scale):
    """roi_align_2d_yx.py:4-7 (the column permutation is what RPOOL_ROI_YX reads in place;
    kept here because this form goes through the Function's xy interface)."""
    return roi_align_2d(x, indices_and_rois[:, [0, 2, 1, 4, 3]], outh, outw, spatial_scale)


class FPNRoIAlign(function.Function):
    """The heads' per-RoI dispatch loop (fpn_roi_mask_head.py:57-63,74-78) as ONE function:
    inputs (indices_and_rois, levels, x_0 .. x_{L-1}), outputs one pooled map per size."""

    def __init__(self, spatial_scales, out_sizes, sampling_ratio=1):
        self.scales = list(spatial_scales)
        self.sizes = [(s, s) if isinstance(s, int) else tuple(s) for s in out_sizes]
        self.sampling_ratio = sampling_ratio

    def forward_gpu(self, inputs):
        rois, levels = inputs[0], inputs[1]
        feats = inputs[2:]
        self.retain_inputs((0, 1))
        self._shapes = [f.shape for f in feats]
        channels = feats[0].shape[1]
        self._rois = cupy.ascontiguousarray(rois)
        self._levels = cupy.ascontiguousarray(levels)
        x = [_nhwc(f) for f in feats]
        tops = [cupy.empty((rois.shape[0], oh, ow, channels), dtype=numpy.float32) for oh, ow in self.sizes]
        p = _problem(x, self.scales, channels, self._rois, _lib.ROI_YX, self._levels, self.sizes, tops,
                     self.sampling_ratio)
        self._ws = _workspace(p)
        L = _lib.lib()
        _lib.check(L.rpool_plan(ctypes.byref(p), self._ws.data.ptr, self._ws.size, _stream()))
        _lib.check(L.rpool_forward(ctypes.byref(p), self._ws.data.ptr, self._ws.size, _stream()))
        return tuple(_nchw(t) for t in tops)

    def backward_gpu(self, inputs, gys):
        channels = self._shapes[0][1]
        diffs = [cupy.empty((s[0], s[2], s[3], s[1]), dtype=numpy.float32) for s in self._shapes]
        g = [_nhwc(gy) for gy in gys]
        p = _problem(diffs, self.scales, channels, self._rois, _lib.ROI_YX, self._levels, self.sizes, g,
                     self.sampling_ratio)
        # the plan of forward_gpu is still in the workspace (same RoIs, levels, geometry)
        _lib.check(_lib.lib().rpool_backward(ctypes.byref(p), self._ws.data.ptr, self._ws.size, _stream()))
        return (None, None) + tuple(_nchw(d) for d in diffs)


def fpn_roi_align(x, indices_and_rois, levels, spatial_scales, out_sizes, sampling_ratio=1):
    """pool_box, pool_mask = fpn_roi_align(x, indices_and_rois, levels, spatial_scales, [7, 14])"""
    outs = FPNRoIAlign(spatial_scales[:len(x)], out_sizes, sampling_ratio)(indices_and_rois, levels, *x)
    return outs
