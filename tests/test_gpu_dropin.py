"""GPU: the two zero/low-edit integration points for the unmodified reference --
the `caffe2_roi_align` module shim and the chainer.Function subclass over CuPy arrays."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

import oracle
import synth
from oracle import reference_loader

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "chainer-maskrcnn_b200", "dropin")


def _reference_available():
    if reference_loader.available():
        return True
    from baseline import refnumpy
    return refnumpy.available()          # points the loader at baseline/_ref


def _case(seed=0, N=2, C=12, H=20, W=27, R=40, scale=0.5):
    rng = np.random.RandomState(seed)
    x = rng.standard_normal((N, C, H, W)).astype(np.float32)
    wh = rng.uniform(2, 30, (R, 2))
    xy = rng.uniform(0, 1, (R, 2)) * (np.array([W, H]) / scale - wh)
    rois = np.concatenate([rng.randint(0, N, (R, 1)), xy, xy + wh], axis=1).astype(np.float32)
    return rng, x, rois, scale


@pytest.mark.skipif(not _reference_available(), reason="no copy of the reference op (baseline/_ref)")
def test_unmodified_reference_forward_cpu_runs_on_the_gpu_through_the_module_shim():
    """roi_align_2d.py:39-46: forward_cpu imports `caffe2_roi_align` and, when that works,
    returns its result.  With chainer-maskrcnn_b200/dropin on sys.path the UNMODIFIED
    reference class therefore pools on the B200; the result must be the C++ port's
    (caffe2 semantics, sampling_ratio 1: caffe2_roi_align.cpp:231-243)."""
    from chainer_maskrcnn_b200 import _lib
    rng, x, rois, scale = _case()
    mod = reference_loader.load_reference_op()
    blocked = sys.modules.get("caffe2_roi_align", "absent")
    sys.path.insert(0, DROPIN)
    sys.modules.pop("caffe2_roi_align", None)
    try:
        shim = importlib.import_module("caffe2_roi_align")
        assert os.path.dirname(shim.__file__) == DROPIN
        n0 = _lib.launch_count()
        f = mod.ROIAlign2D(7, 5, scale)
        (top,) = f.forward_cpu((x, rois))                    # the reference's own method, unmodified
        assert _lib.launch_count() > n0                      # ... ran our kernels
        (top2,) = f.forward_cpu2((x, rois))
    finally:
        sys.path.remove(DROPIN)
        if blocked == "absent":
            sys.modules.pop("caffe2_roi_align", None)
        else:
            sys.modules["caffe2_roi_align"] = blocked
    want = oracle.forward_caffe2(x, rois, 7, 5, scale, 1)
    assert isinstance(top, np.ndarray) and top.dtype == np.float32 and top.shape == want.shape
    assert top.flags["C_CONTIGUOUS"]
    assert oracle.rel_err(top, want) <= 1e-5 and np.array_equal(top, top2)
    assert f._bottom_data_shape == x.shape                   # set by the reference before the import (:40)
    with pytest.raises(RuntimeError):
        shim.forward(x, rois[:, :4], 7, 5, scale)            # "invalid roi shape", caffe2_roi_align.cpp:130-132


def test_chainer_function_subclass_over_cupy_arrays():
    """chainer_adapter.ROIAlign2D: reference base class and conventions (roi_align_2d.py:15-20,
    88-98,190-195), CuPy arrays by address, no torch in the adapter -- against stand-ins for
    chainer and cupy (tests/stub_chainer.py)."""
    import stub_chainer
    remove = stub_chainer.install()
    try:
        ad = importlib.import_module("chainer_maskrcnn_b200.chainer_adapter")
        import cupy
        from chainer import function
        from chainer.utils import type_check
        assert issubclass(ad.ROIAlign2D, function.Function)
        rng, x, rois, scale = _case(seed=1, C=16)
        f = ad.ROIAlign2D(7, 7, scale)
        y = f(cupy.asarray(x), cupy.asarray(rois))
        assert f._retained == (1,) and f._bottom_data_shape == x.shape
        want = oracle.forward_chainer(x, rois, 7, 7, scale)
        assert y.shape == want.shape and oracle.rel_err(cupy.asnumpy(y), want) <= 1e-5
        gy = rng.uniform(-1, 1, want.shape).astype(np.float32)
        gx, none = f.backward_from(cupy.asarray(gy))          # inputs[0] is None here
        assert none is None
        assert oracle.rel_err(cupy.asnumpy(gx), oracle.backward_chainer(gy, rois, x.shape, scale)) <= 1e-4
        # host arrays and the sampling_ratio extension
        (y2,) = ad.ROIAlign2D(7, 7, scale, sampling_ratio=2).forward_cpu((x, rois))
        assert oracle.rel_err(y2, oracle.forward_caffe2(x, rois, 7, 7, scale, 2)) <= 1e-5
        with pytest.raises(type_check.InvalidType):
            f(cupy.asarray(x), cupy.asarray(rois[:, :4]))
        with pytest.raises(type_check.InvalidType):
            f(cupy.asarray(x.astype(np.float64)), cupy.asarray(rois))
        # yx wrapper and the fused head-level function
        yx = oracle.roi_yx_to_xy(rois)                        # the permutation is its own inverse
        y3 = ad._roi_align_2d_yx(cupy.asarray(x), cupy.asarray(yx), 7, 7, scale)
        assert np.array_equal(cupy.asnumpy(y3), cupy.asnumpy(y))
        rng2 = np.random.RandomState(3)
        feats = synth.make_pyramid(rng2, 2, 16, 128, 160, 4)
        r = synth.make_rois(rng2, 2, 50, 128, 160, size_range=(8.0, 150.0))
        lv = oracle.map_rois_to_fpn_levels(r[:, 1:])          # float32, may exceed the pyramid: clipped
        scales = [1.0 / s for s in synth.STRIDES[:4]]
        fn = ad.FPNRoIAlign(scales, [7, 14])
        box, mask = fn(cupy.asarray(r), cupy.asarray(lv), *[cupy.asarray(a) for a in feats])
        lvi = oracle.levels_for_pyramid(r[:, 1:], 4)
        assert oracle.rel_err(cupy.asnumpy(box), oracle.fpn_forward(feats, r, lvi, scales, 7)) <= 1e-5
        assert oracle.rel_err(cupy.asnumpy(mask), oracle.fpn_forward(feats, r, lvi, scales, 14)) <= 1e-5
        g7, g14 = synth.make_gy(rng2, 100, 16, 7), synth.make_gy(rng2, 100, 16, 14)
        grads = fn.backward_from(cupy.asarray(g7), cupy.asarray(g14))
        assert grads[0] is None and grads[1] is None and len(grads) == 6
        a = oracle.fpn_backward(g7, [f_.shape for f_ in feats], r, lvi, scales)
        b = oracle.fpn_backward(g14, [f_.shape for f_ in feats], r, lvi, scales)
        for l in range(4):
            assert oracle.rel_err(cupy.asnumpy(grads[2 + l]), a[l] + b[l]) <= 1e-4
        assert "torch" not in ad.__dict__                     # the adapter binds the C ABI only
    finally:
        remove()
