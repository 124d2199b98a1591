"""ctypes front-end of baseline/refgpu_baseline.cu.  TEST/BENCH INFRASTRUCTURE ONLY.

The reference's CuPy kernels (roi_align_2d.py:100-144, :196-279) restated in
CUDA and driven per RoI like the reference's FPN heads
(fpn_roi_mask_head.py:57-63): the same-box GPU baseline of bench.py
("gpu_baseline") and a cross-check in tests/.  Never imported by the product.
"""
import ctypes
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "refgpu_baseline.cu")
_LIB_PATH = os.path.join(_HERE, "librefgpu_baseline.so")
_lib = None


class Level(ctypes.Structure):
    _fields_ = [("x", ctypes.c_void_p), ("grad", ctypes.c_void_p), ("tmp", ctypes.c_void_p),
                ("N", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int),
                ("scale", ctypes.c_float)]


def build(force=False):
    stale = not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(_SRC)
    if not (force or stale):
        return _LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
                           "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-o", _LIB_PATH, _SRC])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, i, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        L.refgpu_forward.argtypes = [vp, i, i, i, vp, i, i, i, f, vp, vp]
        L.refgpu_backward.argtypes = [vp, vp, i, i, i, i, i, i, i, f, vp, vp]
        L.refgpu_fpn_step.argtypes = [ctypes.POINTER(Level), i, i, vp, ctypes.POINTER(ctypes.c_int), i, i,
                                      vp, vp, i, vp]
        L.refgpu_fpn_step.restype = ctypes.c_longlong
        _lib = L
    return _lib


def _stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def forward(x, rois_xy, outh, outw, scale):
    """x: (N,C,H,W) contiguous CUDA float32; rois_xy: (R,5) [b,x1,y1,x2,y2]."""
    import torch
    n, c, h, w = x.shape
    top = torch.empty((rois_xy.shape[0], c, outh, outw), dtype=torch.float32, device=x.device)
    rc = lib().refgpu_forward(x.data_ptr(), c, h, w, rois_xy.data_ptr(), rois_xy.shape[0], outh, outw,
                              float(scale), top.data_ptr(), _stream())
    assert rc == 0, rc
    return top


def backward(gy, rois_xy, shape, scale):
    import torch
    n, c, h, w = shape
    gx = torch.empty(shape, dtype=torch.float32, device=gy.device)
    rc = lib().refgpu_backward(gy.data_ptr(), rois_xy.data_ptr(), rois_xy.shape[0], n, c, h, w,
                               gy.shape[2], gy.shape[3], float(scale), gx.data_ptr(), _stream())
    assert rc == 0, rc
    return gx


class FpnState(object):
    """Device buffers of the per-RoI dispatch: NCHW features, gradients and the
    per-call dense gradient of every level."""

    def __init__(self, feats_nchw, scales):
        import torch
        self.feats = [f.contiguous() for f in feats_nchw]
        self.grads = [torch.empty_like(f) for f in self.feats]
        self.tmps = [torch.empty_like(f) for f in self.feats]
        self.levels = (Level * len(self.feats))()
        for l, (f, g, t, s) in enumerate(zip(self.feats, self.grads, self.tmps, scales)):
            self.levels[l].x, self.levels[l].grad, self.levels[l].tmp = f.data_ptr(), g.data_ptr(), t.data_ptr()
            self.levels[l].N, self.levels[l].H, self.levels[l].W = f.shape[0], f.shape[2], f.shape[3]
            self.levels[l].scale = float(s)
        self.C = self.feats[0].shape[1]


def fpn_step(state, rois_xy, levels_host, P, top, gy, do_backward=True):
    """One fwd(+bwd) pass of the heads' per-RoI loop; returns the number of device
    operations (kernel launches + memsets/copies) it enqueued."""
    lv = (ctypes.c_int * len(levels_host))(*[int(v) for v in levels_host])
    ops = lib().refgpu_fpn_step(state.levels, len(state.feats), state.C, rois_xy.data_ptr(), lv,
                                rois_xy.shape[0], P, top.data_ptr(), gy.data_ptr() if gy is not None else None,
                                int(bool(do_backward)), _stream())
    assert ops >= 0
    return int(ops)
