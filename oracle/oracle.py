"""ctypes/NumPy front-end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it restates (paths relative to the
reference repository root).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_roialign.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libref_caffe2_roialign.so")

__all__ = [
    "build", "have_ref", "max_threads",
    "forward_chainer", "backward_chainer", "forward_caffe2", "backward_caffe2",
    "ref_caffe2_forward", "roi_yx_to_xy",
    "map_rois_to_fpn_levels", "level_area_thresholds", "levels_for_pyramid",
    "fpn_forward", "fpn_backward", "rel_err", "err_stats",
]

_lib = None
_ref = None

_f32p = ctypes.POINTER(ctypes.c_float)


def build(force=False):
    """Compile oracle/liboracle_roialign.so (and oracle/_ref when the
    reference tree is present).  Building the checker is not using it."""
    src = os.path.join(_HERE, "roialign_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    ref_src = "/root/reference/chainer_maskrcnn/functions/roi_align/caffe2_operation/caffe2_roi_align.cpp"
    if os.path.exists(ref_src) and (force or not os.path.exists(_REF_PATH)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        i = ctypes.c_int
        f = ctypes.c_float
        _lib.orc_forward_chainer_mt.argtypes = [_f32p, i, i, i, i, _f32p, i, i, i, f, _f32p, i]
        _lib.orc_backward_chainer_mt.argtypes = [_f32p, _f32p, i, i, i, i, i, i, i, f, _f32p, i]
        _lib.orc_forward_caffe2_mt.argtypes = [_f32p, i, i, i, i, _f32p, i, i, i, f, i, _f32p, i]
        _lib.orc_backward_caffe2_mt.argtypes = [_f32p, _f32p, i, i, i, i, i, i, i, f, i, _f32p, i]
        for fn in (_lib.orc_forward_chainer_mt, _lib.orc_backward_chainer_mt,
                   _lib.orc_forward_caffe2_mt, _lib.orc_backward_caffe2_mt,
                   _lib.orc_max_threads):
            fn.restype = i
    return _lib


def have_ref():
    """True when the reference's own C++ forward was compiled into oracle/_ref."""
    return os.path.exists(_REF_PATH)


def _load_ref():
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(_REF_PATH)
        i = ctypes.c_int
        _ref.ref_caffe2_roi_align_forward.argtypes = [
            _f32p, i, i, i, i, _f32p, i, i, i, ctypes.c_float, i, _f32p]
        _ref.ref_caffe2_roi_align_forward.restype = i
    return _ref


def max_threads():
    return int(_load().orc_max_threads())


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _check(err, what):
    if err:
        # the reference raises IndexError for taps beyond the map
        raise IndexError("%s: RoI %d samples outside the feature map" % (what, err - 1))


def forward_chainer(x, rois, outh, outw, spatial_scale, threads=1):
    """ROIAlign2D.forward_cpu, NumPy body (roi_align_2d.py:48-88).
    x (N,C,H,W) f32, rois (R,5) [b,x1,y1,x2,y2] f32 -> (R,C,outh,outw) f32."""
    x, xp = _c(x)
    rois, rp = _c(rois)
    N, C, H, W = x.shape
    R = rois.shape[0]
    top = np.zeros((R, C, outh, outw), np.float32)
    err = _load().orc_forward_chainer_mt(xp, N, C, H, W, rp, R, outh, outw,
                                         np.float32(spatial_scale), top.ctypes.data_as(_f32p),
                                         threads)
    _check(err, "forward_chainer")
    return top


def backward_chainer(gy, rois, x_shape, spatial_scale, threads=1):
    """ROIAlign2D.backward_cpu (roi_align_2d.py:148-190) -> dense (N,C,H,W) f32."""
    gy, gp = _c(gy)
    rois, rp = _c(rois)
    N, C, H, W = x_shape
    R, C2, outh, outw = gy.shape
    assert C2 == C
    bottom = np.zeros((N, C, H, W), np.float32)
    err = _load().orc_backward_chainer_mt(gp, rp, R, N, C, H, W, outh, outw,
                                          np.float32(spatial_scale),
                                          bottom.ctypes.data_as(_f32p), threads)
    _check(err, "backward_chainer")
    return bottom


def forward_caffe2(x, rois, outh, outw, spatial_scale, sampling_ratio, threads=1):
    """Restatement of ROIAlignForward<float> (caffe2_roi_align.cpp:115-226)."""
    x, xp = _c(x)
    rois, rp = _c(rois)
    N, C, H, W = x.shape
    R = rois.shape[0]
    top = np.zeros((R, C, outh, outw), np.float32)
    err = _load().orc_forward_caffe2_mt(xp, N, C, H, W, rp, R, outh, outw,
                                        np.float32(spatial_scale), int(sampling_ratio),
                                        top.ctypes.data_as(_f32p), threads)
    _check(err, "forward_caffe2")
    return top


def backward_caffe2(gy, rois, x_shape, spatial_scale, sampling_ratio, threads=1):
    """Exact adjoint of forward_caffe2 (the reference ships none)."""
    gy, gp = _c(gy)
    rois, rp = _c(rois)
    N, C, H, W = x_shape
    R, C2, outh, outw = gy.shape
    assert C2 == C
    bottom = np.zeros((N, C, H, W), np.float32)
    err = _load().orc_backward_caffe2_mt(gp, rp, R, N, C, H, W, outh, outw,
                                         np.float32(spatial_scale), int(sampling_ratio),
                                         bottom.ctypes.data_as(_f32p), threads)
    _check(err, "backward_caffe2")
    return bottom


def ref_caffe2_forward(x, rois, outh, outw, spatial_scale, sampling_ratio):
    """The reference's own compiled C++ (oracle/_ref), single-threaded as shipped
    (its omp pragma is commented out, caffe2_roi_align.cpp:137-138)."""
    x, xp = _c(x)
    rois, rp = _c(rois)
    N, C, H, W = x.shape
    R = rois.shape[0]
    top = np.zeros((R, C, outh, outw), np.float32)  # binding zero-fills, :237-238
    err = _load_ref().ref_caffe2_roi_align_forward(
        xp, N, C, H, W, rp, R, outh, outw, np.float32(spatial_scale), int(sampling_ratio),
        top.ctypes.data_as(_f32p))
    if err:
        raise RuntimeError("reference C++ forward raised")
    return top


def roi_yx_to_xy(indices_and_rois):
    """_roi_align_2d_yx's column permutation (roi_align_2d_yx.py:5)."""
    return np.ascontiguousarray(np.asarray(indices_and_rois)[:, [0, 2, 1, 4, 3]])


# ---------------------------------------------------------------------------
# level assignment
# ---------------------------------------------------------------------------

def map_rois_to_fpn_levels(rois, k_min=0, k_max=4, s0=224, lvl0=4, eps=1e-6):
    """map_rois_to_fpn_levels (model/rpn/multilevel_region_proposal_network.py:16-31),
    the same NumPy expression on (R,4) [y1,x1,y2,x2] float32; returns float32."""
    rois = np.asarray(rois, dtype=np.float32)
    area = np.prod(rois[:, 2:] - rois[:, :2], axis=1)
    s = np.sqrt(area)
    target = np.floor(lvl0 + np.log2(s / s0 + eps))
    return np.clip(target, k_min, k_max)


def levels_for_pyramid(rois, n_levels, k_min=0, k_max=4):
    """Level mapper followed by MaskRCNN.__call__'s clip to the pyramid
    (model/maskrcnn.py:141) and the head's int32 cast (fpn_roi_mask_head.py:58)."""
    lv = map_rois_to_fpn_levels(rois, k_min, k_max)
    return np.clip(lv, 0, n_levels - 1).astype(np.int32)


def _level_of_area(area_f32, s0, lvl0, eps):
    a = np.array([area_f32], dtype=np.float32)
    with np.errstate(divide="ignore"):
        return float(np.floor(lvl0 + np.log2(np.sqrt(a) / s0 + eps))[0])


def level_area_thresholds(k_min=0, k_max=4, s0=224, lvl0=4, eps=1e-6):
    """Smallest float32 area mapped to level >= k, for k = k_min+1 .. k_max, found
    by bisection over float32 bit patterns against the NumPy expression above
    (SURVEY.md appendix A.2).  level == k_min + #{thresholds <= area}."""
    out = []
    for k in range(k_min + 1, k_max + 1):
        lo = np.float32(0.0).view(np.uint32).item()        # level(lo) < k
        hi = np.float32(3.0e38).view(np.uint32).item()     # level(hi) >= k
        assert _level_of_area(np.uint32(hi).view(np.float32), s0, lvl0, eps) >= k
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if _level_of_area(np.uint32(mid).view(np.float32), s0, lvl0, eps) >= k:
                hi = mid
            else:
                lo = mid
        out.append(np.uint32(hi).view(np.float32))
    return np.array(out, dtype=np.float32)


# ---------------------------------------------------------------------------
# FPN heads' dispatch loop
# ---------------------------------------------------------------------------

def _single(mode, sampling_ratio):
    if mode == "chainer":
        assert sampling_ratio == 1
        return (lambda x, r, oh, ow, s, t: forward_chainer(x, r, oh, ow, s, t),
                lambda g, r, shp, s, t: backward_chainer(g, r, shp, s, t))
    assert mode == "caffe2"
    return (lambda x, r, oh, ow, s, t: forward_caffe2(x, r, oh, ow, s, sampling_ratio, t),
            lambda g, r, shp, s, t: backward_caffe2(g, r, shp, s, sampling_ratio, t))


def fpn_forward(features, indices_and_rois, levels, spatial_scales, out_size,
                mode="chainer", sampling_ratio=1, threads=1):
    """FPNRoIMaskHead's pooling loop (model/head/fpn_roi_mask_head.py:57-63):
    RoI r (columns [idx,y1,x1,y2,x2]) is pooled from features[levels[r]] with
    spatial_scales[levels[r]] and the results are concatenated in input order.
    Evaluated one call per level (same values, RoIs are independent)."""
    fwd, _ = _single(mode, sampling_ratio)
    rois_xy = roi_yx_to_xy(np.asarray(indices_and_rois, np.float32))
    levels = np.asarray(levels).astype(np.int32)
    R = rois_xy.shape[0]
    C = features[0].shape[1]
    oh, ow = out_size if isinstance(out_size, (tuple, list)) else (out_size, out_size)
    out = np.zeros((R, C, oh, ow), np.float32)
    for l in range(len(features)):
        sel = np.nonzero(levels == l)[0]
        if sel.size:
            out[sel] = fwd(features[l], rois_xy[sel], oh, ow, spatial_scales[l], threads)
    return out


def fpn_backward(gy, feature_shapes, indices_and_rois, levels, spatial_scales,
                 mode="chainer", sampling_ratio=1, threads=1):
    """Gradient of fpn_forward w.r.t. each level's feature map: per level, the
    sum over that level's RoIs of the op's backward (what Chainer's autograd
    accumulates over the per-RoI FunctionNodes), RoIs visited in input order."""
    _, bwd = _single(mode, sampling_ratio)
    rois_xy = roi_yx_to_xy(np.asarray(indices_and_rois, np.float32))
    levels = np.asarray(levels).astype(np.int32)
    grads = []
    for l, shp in enumerate(feature_shapes):
        sel = np.nonzero(levels == l)[0]
        if sel.size:
            grads.append(bwd(np.ascontiguousarray(gy[sel]), rois_xy[sel], tuple(shp),
                             spatial_scales[l], threads))
        else:
            grads.append(np.zeros(tuple(shp), np.float32))
    return grads


def rel_err(a, b):
    """The parity metric of SURVEY.md 8(c): max|a-b| / max|b| (b = oracle)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.max(np.abs(b)) if b.size else 0.0
    if denom == 0.0:
        return float(np.max(np.abs(a))) if a.size else 0.0
    return float(np.max(np.abs(a - b)) / denom)


def err_stats(a, b):
    """Achieved errors of `a` against the oracle `b`, both readings of "relative error":
      max_norm  max|a-b| / max|b|                       (rel_err, the gate of the parity tests)
      elem_rel  max_i |a_i-b_i| / max(|b_i|, rms(b))    (element-wise, floored at the RMS
                magnitude so that entries near zero do not divide by nothing)
      max_abs   max|a-b|
    """
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if b.size == 0:
        return {"max_norm": 0.0, "elem_rel": 0.0, "max_abs": 0.0}
    d = np.abs(a - b)
    rms = float(np.sqrt(np.mean(b * b)))
    floor = rms if rms > 0 else 1.0
    return {"max_norm": rel_err(a, b), "elem_rel": float(np.max(d / np.maximum(np.abs(b), floor))),
            "max_abs": float(d.max())}
