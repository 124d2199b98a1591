"""Host-side engine: turns (pyramid, RoIs, levels) into rpool_problem blocks and
launches the library on torch's current stream.  torch is used for device
memory, streams and autograd plumbing only -- all arithmetic on the path runs in
librpool_b200.so.

Layout contract: the kernels read channels-last features and write channels-last
pooled maps.  Logical shapes stay the reference's (N,C,H,W) / (R,C,PH,PW):
tensors are returned in ``torch.channels_last`` memory format, and NCHW-contiguous
inputs are converted once per call by the library's own transpose kernel.
"""
import ctypes
import functools

import numpy as np
import torch

from . import _lib


# ---------------------------------------------------------------------------
# level thresholds: the reference's NumPy arithmetic, tabulated
# ---------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def level_thresholds(k_min=0, k_max=4, s0=224, lvl0=4, eps=1e-6):
    """float32 areas at which floor(lvl0 + log2(sqrt(area)/s0 + eps)) -- evaluated
    exactly like map_rois_to_fpn_levels
    (chainer_maskrcnn/model/rpn/multilevel_region_proposal_network.py:24-30, NumPy
    float32) -- first reaches k_min+1 .. k_max.  On the device the level is then
    k_min + #{thresholds <= area}: IEEE sub/mul/compare only, hence bit-exact."""
    def level_of(bits):
        a = np.array([bits], dtype=np.uint32).view(np.float32)
        with np.errstate(divide="ignore"):
            return float(np.floor(lvl0 + np.log2(np.sqrt(a) / s0 + eps))[0])
    out = []
    for k in range(k_min + 1, k_max + 1):
        lo, hi = 0, 0x7F000000
        if not level_of(hi) >= k:
            raise ValueError("level %d is unreachable" % k)
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if level_of(mid) >= k:
                hi = mid
            else:
                lo = mid
        out.append(float(np.array([hi], dtype=np.uint32).view(np.float32)[0]))
    if any(b < a for a, b in zip(out, out[1:])):
        raise AssertionError("level thresholds are not monotone: %r" % (out,))
    return tuple(out)


# ---------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------
_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """Handle of torch's current stream on the current device.  The raw accessor is
    used when torch has it: torch.cuda.current_stream() builds a Stream object through
    several Python layers (14 us per call here; three calls per forward + backward)."""
    if _RAW_STREAM is not None:
        return ctypes.c_void_p(_RAW_STREAM(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _NoGuard(object):
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def _on(device):
    """Device guard for the launches: nothing when `device` is already current (the
    usual case; the torch guard costs several microseconds per call)."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)


def _require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA torch.Tensor (there is no CPU fallback)" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (name, t.dtype))


def to_channels_last(t):
    """Logical (N,C,H,W) float32 CUDA tensor -> same values, physically NHWC.
    Zero-copy when it already is; otherwise one pass of rpool_nchw_to_nhwc."""
    if t.is_contiguous(memory_format=torch.channels_last):
        return t
    if t.numel() == 0:
        return t.contiguous(memory_format=torch.channels_last)
    src = t.contiguous()
    n, c, h, w = src.shape
    dst = torch.empty_like(src, memory_format=torch.channels_last)
    with _on(t.device):     # launch on the tensor's device and its current stream
        _lib.check(_lib.lib().rpool_nchw_to_nhwc(src.data_ptr(), dst.data_ptr(), n, c, h, w, _stream()))
    return dst


def _pad_channels(t, cpad):
    """Logical (N,C,H,W) CUDA tensor -> channels-last (N,cpad,H,W), zeros in channels >= C."""
    n, c, h, w = t.shape
    out = torch.empty((n, cpad, h, w), dtype=t.dtype, device=t.device, memory_format=torch.channels_last)
    if out.numel():
        out[:, :c].copy_(t)
        out[:, c:].zero_()
    return out


def to_nchw_contiguous(t):
    """Logical (N,C,H,W) tensor in channels-last memory -> NCHW-contiguous copy."""
    if t.is_contiguous() or t.numel() == 0:
        return t.contiguous()
    if not t.is_contiguous(memory_format=torch.channels_last):
        t = to_channels_last(t)
    n, c, h, w = t.shape
    dst = torch.empty((n, c, h, w), dtype=t.dtype, device=t.device)
    with _on(t.device):
        _lib.check(_lib.lib().rpool_nhwc_to_nchw(t.data_ptr(), dst.data_ptr(), n, c, h, w, _stream()))
    return dst


class Plan(object):
    """Everything backward needs: geometry, RoIs, the device schedule."""
    __slots__ = ("shapes", "scales", "rois", "levels_i32", "levels_f32", "thresholds", "k_min",
                 "out_sizes", "sampling_ratio", "coord_mode", "roi_format", "workspace",
                 "channels", "cpad", "device", "problem", "options")


def _fill_problem(plan, level_ptrs, pooled_ptrs, accumulate=False, deterministic=False):
    """The plan's rpool_problem block with this call's addresses patched in.  The
    geometry is written once per plan; ctypes field stores are the bulk of the host
    time of a small call."""
    p = plan.problem
    if p is not None:
        for l, ptr in enumerate(level_ptrs):
            p.level[l].data = ptr
        for h, ptr in enumerate(pooled_ptrs):
            p.pooled[h] = ptr
        p.accumulate = int(accumulate)
        p.deterministic = int(bool(deterministic))
        p.det_workspace = None
        p.det_workspace_bytes = 0
        return p
    p = plan.problem = _lib.Problem()
    p.n_levels = len(plan.shapes)
    p.channels = plan.cpad
    p.feat_layout = _lib.NHWC
    p.pool_layout = _lib.NHWC
    for l, (shape, scale, ptr) in enumerate(zip(plan.shapes, plan.scales, level_ptrs)):
        p.level[l].data = ptr
        p.level[l].n_images = shape[0]
        p.level[l].height = shape[2]
        p.level[l].width = shape[3]
        p.level[l].spatial_scale = float(scale)
    p.rois = plan.rois.data_ptr()
    p.n_rois = plan.rois.shape[0]
    p.roi_format = plan.roi_format
    p.roi_levels = plan.levels_i32.data_ptr() if plan.levels_i32 is not None else None
    p.roi_levels_f32 = plan.levels_f32.data_ptr() if plan.levels_f32 is not None else None
    for t, v in enumerate(plan.thresholds):
        p.level_thresholds[t] = v
    p.n_thresholds = len(plan.thresholds)
    p.k_min = plan.k_min
    p.n_heads = len(plan.out_sizes)
    for h, ((oh, ow), ptr) in enumerate(zip(plan.out_sizes, pooled_ptrs)):
        p.out_h[h] = oh
        p.out_w[h] = ow
        p.pooled[h] = ptr
    p.sampling_ratio = plan.sampling_ratio
    p.coord_mode = plan.coord_mode
    p.accumulate = int(accumulate)
    p.deterministic = int(bool(deterministic))
    p.opt = _lib.make_options(**plan.options)
    return p


def _norm_sizes(out_sizes):
    sizes = []
    for s in out_sizes:
        if isinstance(s, (tuple, list)):
            sizes.append((int(s[0]), int(s[1])))
        else:
            sizes.append((int(s), int(s)))
    if not 1 <= len(sizes) <= _lib.MAX_HEADS:
        raise ValueError("between 1 and %d pooled sizes per call, got %d" % (_lib.MAX_HEADS, len(sizes)))
    return sizes


def default_coord_mode(sampling_ratio):
    return _lib.COORD_CHAINER if sampling_ratio == 1 else _lib.COORD_CAFFE2


def make_plan(shapes, rois, levels, spatial_scales, out_sizes, sampling_ratio=1,
              coord_mode=None, roi_format=_lib.ROI_YX, k_min=0, k_max=4, pad_channels=True,
              options=None, workspace=None):
    """Validates the call, assigns levels (when not given) and bins the RoIs by
    (image, level) on the device: one small launch (rpool_plan).  The Plan is
    all that forward and backward share.  ``options``: rpool_options fields by name
    (experiments; defaults otherwise).  ``workspace``: a uint8 CUDA tensor to reuse."""
    shapes = [tuple(int(v) for v in s) for s in shapes]
    if not 1 <= len(shapes) <= _lib.MAX_LEVELS:
        raise ValueError("pyramid must have 1..%d levels" % _lib.MAX_LEVELS)
    if len(spatial_scales) != len(shapes):
        raise ValueError("one spatial_scale per level")
    if any(len(s) != 4 for s in shapes):
        raise TypeError("features must be 4-D (N,C,H,W)")
    _require_cuda(rois, "rois")
    if rois.dim() != 2 or rois.shape[1] != 5:
        raise TypeError("rois must have shape (R,5), got %s" % (tuple(rois.shape),))
    C = shapes[0][1]
    if any(s[1] != C for s in shapes):
        raise ValueError("all levels must have the same channel count")
    if coord_mode is None:
        coord_mode = default_coord_mode(sampling_ratio)

    plan = Plan()
    plan.problem = None
    plan.options = dict(options or {})
    plan.device = rois.device
    plan.shapes = shapes
    plan.scales = [float(s) for s in spatial_scales]
    plan.rois = rois.contiguous()
    plan.levels_i32 = plan.levels_f32 = None
    plan.thresholds = ()
    plan.k_min = int(k_min)
    if levels is not None:
        if not isinstance(levels, torch.Tensor) or not levels.is_cuda:
            raise TypeError("levels must be a CUDA tensor or None")
        if tuple(levels.shape) != (rois.shape[0],):
            raise ValueError("levels must have shape (R,)")
        if levels.dtype == torch.float32:
            plan.levels_f32 = levels.contiguous()       # map_rois_to_fpn_levels' own dtype
        elif levels.dtype in (torch.float64, torch.float16, torch.bfloat16):
            plan.levels_f32 = levels.to(torch.float32).contiguous()
        elif levels.dtype == torch.bool or levels.dtype.is_complex:
            raise TypeError("levels must be an integer or floating tensor, got %s" % levels.dtype)
        else:
            # the heads cast with astype(int32) (fpn_roi_mask_head.py:58): any integer dtype will do
            plan.levels_i32 = levels.to(torch.int32).contiguous()
    elif len(shapes) > 1:
        plan.thresholds = level_thresholds(k_min, k_max)
    plan.out_sizes = _norm_sizes(out_sizes)
    plan.sampling_ratio = int(sampling_ratio)
    plan.coord_mode = int(coord_mode)
    plan.roi_format = int(roi_format)
    plan.channels = int(C)
    # the vectorised kernel path moves 4 channels per lane: other channel counts (the 490-channel
    # thin map of light_roi_mask_head.py:26) are padded with zero channels on the way in and
    # cut on the way out (4x faster than the scalar generic path at C = 490)
    plan.cpad = (int(C) + 3) // 4 * 4 if pad_channels else int(C)

    L = _lib.lib()
    ws_bytes = L.rpool_workspace_bytes_ex(rois.shape[0], len(plan.out_sizes), plan.coord_mode)
    if workspace is not None:
        if workspace.dtype != torch.uint8 or workspace.device != rois.device or workspace.numel() < ws_bytes:
            raise ValueError("workspace must be a uint8 tensor of >= %d bytes on %s" % (ws_bytes, rois.device))
        plan.workspace = workspace
    else:
        plan.workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=rois.device)
    ws_bytes = plan.workspace.numel()
    # rpool_plan reads geometry only; level/pooled addresses are not dereferenced
    dummy = plan.workspace.data_ptr()
    prob = _fill_problem(plan, [dummy] * len(shapes), [None] * len(plan.out_sizes))
    with _on(rois.device):
        _lib.check(L.rpool_plan(ctypes.byref(prob), plan.workspace.data_ptr(), ws_bytes, _stream()))
    return plan


def forward(features, rois, levels, spatial_scales, out_sizes, sampling_ratio=1,
            coord_mode=None, roi_format=_lib.ROI_YX, k_min=0, k_max=4, plan=None, pad_channels=True,
            options=None, workspace=None, out=None):
    """One fused launch over the pyramid.  Returns ([pooled per head], Plan).

    features: sequence of logical (N,C,H_l,W_l) float32 CUDA tensors;
    rois: (R,5) float32 CUDA tensor; levels: None (assigned on the device by the
    reference rule, clipped to the pyramid), or an integer/floating CUDA tensor (R,).
    ``out``: optional preallocated channels-last (R,C,oh,ow) tensors to write into.
    """
    features = list(features)
    for i, f in enumerate(features):
        _require_cuda(f, "features[%d]" % i)
    if plan is None:
        plan = make_plan([f.shape for f in features], rois, levels, spatial_scales, out_sizes,
                         sampling_ratio, coord_mode, roi_format, k_min, k_max, pad_channels,
                         options, workspace)
    R, C, Cp = plan.rois.shape[0], plan.channels, plan.cpad
    for i, f in enumerate(features):
        if f.device != plan.device:
            raise ValueError("features[%d] is on %s, the RoIs are on %s" % (i, f.device, plan.device))
    feats = [to_channels_last(f) if Cp == C else _pad_channels(f, Cp) for f in features]
    if out is not None and Cp == C:
        outs = list(out)
        for o, (oh, ow) in zip(outs, plan.out_sizes):
            if tuple(o.shape) != (R, C, oh, ow) or not o.is_contiguous(memory_format=torch.channels_last) \
                    or o.dtype != torch.float32 or o.device != plan.device:
                raise ValueError("out tensors must be float32 channels-last (R,C,oh,ow) on the RoIs' device")
    else:
        outs = [torch.empty((R, Cp, oh, ow), dtype=torch.float32, device=plan.device,
                            memory_format=torch.channels_last) for oh, ow in plan.out_sizes]
    prob = _fill_problem(plan, [f.data_ptr() for f in feats], [o.data_ptr() for o in outs])
    with _on(plan.device):
        _lib.check(_lib.lib().rpool_forward(ctypes.byref(prob), plan.workspace.data_ptr(),
                                            plan.workspace.numel(), _stream()))
    if Cp != C:
        outs = [o[:, :C] for o in outs]
    return outs, plan


def backward(plan, gys, deterministic=False, out=None, accumulate=False, check_flags=True,
             det_scratch=None, fill_in_tail=False):
    """Dense feature gradients (channels-last memory, logical (N,C,H,W)) for
    every level from the pooled gradients ``gys`` (one per head).  ``out``: optional
    preallocated gradient tensors to write into; with ``accumulate=True`` the result
    is added to what ``out`` already holds (no zero fill: rpool_problem.accumulate).
    ``fill_in_tail``: the zero fill may start while the kernel queued before this call on the
    stream is still draining; only when that kernel does not touch the gradient buffers
    (``rpool_options.zero_fill_in_tail``; FusedStep: it is the forward launch).
    ``deterministic``: segmented reduction instead of atomics (bit-identical from run to
    run).  Its scratch holds one private window per RoI; without ``det_scratch`` (a uint8
    CUDA tensor) the exact size is computed on the device first, which synchronises the
    stream.  RoIs that need the generic kernel path cannot be ordered, and a scratch that
    is too small cannot hold the windows: both are reported -- ``check_flags`` reads the
    flags back (one more synchronisation) and raises; pass a scratch and False inside
    CUDA-graph capture and call ``status_flags(plan)`` afterwards."""
    if accumulate and out is None:
        raise ValueError("accumulate=True needs the gradient tensors to add into (out=...)")
    if len(gys) != len(plan.out_sizes):
        raise ValueError("one gy per head")
    R = plan.rois.shape[0]
    g_in = []
    for h, (g, (oh, ow)) in enumerate(zip(gys, plan.out_sizes)):
        if g is None:
            g = torch.zeros((R, plan.channels, oh, ow), dtype=torch.float32, device=plan.device,
                            memory_format=torch.channels_last)
        _require_cuda(g, "gy[%d]" % h)
        if tuple(g.shape) != (R, plan.channels, oh, ow):
            raise ValueError("gy[%d] has shape %s, expected %s" %
                             (h, tuple(g.shape), (R, plan.channels, oh, ow)))
        g_in.append(to_channels_last(g) if plan.cpad == plan.channels else _pad_channels(g, plan.cpad))
    padded = plan.cpad != plan.channels
    user_out = None
    if out is not None:
        user_out = list(out)
        for g, shape in zip(user_out, plan.shapes):
            if tuple(g.shape) != shape or not g.is_contiguous(memory_format=torch.channels_last):
                raise ValueError("out gradients must be channels-last tensors of the feature shapes")
    if out is None or padded:
        grads = [torch.empty((shape[0], plan.cpad, shape[2], shape[3]), dtype=torch.float32,
                             device=plan.device, memory_format=torch.channels_last)
                 for shape in plan.shapes]
    else:
        grads = user_out
    prob = _fill_problem(plan, [g.data_ptr() for g in grads], [g.data_ptr() for g in g_in],
                         accumulate=accumulate and not padded, deterministic=deterministic)
    prob.opt.zero_fill_in_tail = int(bool(fill_in_tail))
    L = _lib.lib()
    with _on(plan.device):
        ws, ws_n = plan.workspace.data_ptr(), plan.workspace.numel()
        if deterministic:
            if det_scratch is None:
                det_scratch = torch.empty(det_scratch_bytes(plan, prob), dtype=torch.uint8, device=plan.device)
            prob.det_workspace = det_scratch.data_ptr()
            prob.det_workspace_bytes = det_scratch.numel()
        _lib.check(L.rpool_backward(ctypes.byref(prob), ws, ws_n, _stream()))
        if deterministic and check_flags:
            err = ctypes.c_int32(0)
            _lib.check(L.rpool_status_flags(ws, R, _stream(), ctypes.byref(err)))
            if err.value & (_lib.FLAG_DET_GENERIC | _lib.FLAG_DET_SCRATCH):
                raise _lib.RpoolError(_lib.UNSUPPORTED, "deterministic backward: some RoIs need the generic "
                                      "kernel path (pooled size > 16, sampling grid > 4, window taller "
                                      "than 64 rows or map narrower than 8 columns); flags %d" % err.value)
    if padded:
        C = plan.channels
        if user_out is None:
            return [g[:, :C] for g in grads]
        for dst, g in zip(user_out, grads):
            if accumulate:
                dst.add_(g[:, :C])
            else:
                dst.copy_(g[:, :C])
        return user_out
    return grads


def det_scratch_bytes(plan, prob=None):
    """Exact size of the deterministic backward's scratch for this plan's RoIs: computed on
    the device, SYNCHRONISES the current stream (rpool_backward_det_bytes)."""
    if prob is None:
        dummy = plan.workspace.data_ptr()
        prob = _fill_problem(plan, [dummy] * len(plan.shapes), [dummy] * len(plan.out_sizes), deterministic=True)
    need = ctypes.c_size_t(0)
    with _on(plan.device):
        _lib.check(_lib.lib().rpool_backward_det_bytes(ctypes.byref(prob), plan.workspace.data_ptr(),
                                                       plan.workspace.numel(), _stream(), ctypes.byref(need)))
    return need.value


def read_plan(plan):
    """(levels, order) of the device schedule as NumPy int32 arrays (tests)."""
    R = plan.rois.shape[0]
    lv = np.empty(R, np.int32)
    od = np.empty(R, np.int32)
    with torch.cuda.device(plan.device):
        _lib.check(_lib.lib().rpool_read_plan(plan.workspace.data_ptr(), R, lv.ctypes.data,
                                              od.ctypes.data, _stream()))
    return lv, od


def status_flags(plan):
    """RPOOL_FLAG_* bits raised for this plan's RoIs (synchronises the current stream).
    The reference's NumPy path raises IndexError for a RoI whose batch index or taps
    leave the tensor (roi_align_2d.py:76-86); the kernels compute zeros / clamp instead
    and flag it here."""
    flags = ctypes.c_int32(0)
    with _on(plan.device):
        _lib.check(_lib.lib().rpool_status_flags(plan.workspace.data_ptr(), plan.rois.shape[0],
                                                 _stream(), ctypes.byref(flags)))
    return flags.value


def check_rois(plan):
    """Raises IndexError, like the reference's NumPy path, when a RoI addressed an image
    outside the batch (synchronises the current stream)."""
    if status_flags(plan) & _lib.FLAG_BAD_BATCH:
        raise IndexError("a RoI's batch index is outside [0, n_images)")


def zero_fill(grads):
    """Zero fill of channels-last gradient tensors (one launch for all levels) on the
    current stream: rpool_backward's first step, callable early on another stream."""
    grads = list(grads)
    if not grads:
        return
    p = _lib.Problem()
    p.n_levels = len(grads)
    p.channels = grads[0].shape[1]
    for l, g in enumerate(grads):
        _require_cuda(g, "grads[%d]" % l)
        if not g.is_contiguous(memory_format=torch.channels_last) or g.shape[1] != p.channels:
            raise ValueError("gradients must be channels-last tensors with one channel count")
        p.level[l].data = g.data_ptr()
        p.level[l].n_images, p.level[l].height, p.level[l].width = g.shape[0], g.shape[2], g.shape[3]
    with _on(grads[0].device):
        _lib.check(_lib.lib().rpool_zero_fill(ctypes.byref(p), _stream()))


class _FPNRoIAlignFn(torch.autograd.Function):
    """autograd node: the analogue of the reference's ROIAlign2D Function objects
    (roi_align_2d.py:15), one per call instead of one per RoI."""

    @staticmethod
    def forward(ctx, rois, levels, cfg, *features):
        outs, plan = forward(features, rois, levels, **cfg)
        ctx.plan = plan
        return tuple(outs)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gys):
        grads = backward(ctx.plan, list(gys))
        return (None, None, None) + tuple(grads)


def apply(features, rois, levels, **cfg):
    """Differentiable (w.r.t. the features) fused pooling; tuple of pooled maps."""
    return _FPNRoIAlignFn.apply(rois, levels, cfg, *features)


def assign_levels(boxes, roi_format=_lib.ROI_YX, k_min=0, k_max=4, k_cap=None, as_int=False):
    """Device map_rois_to_fpn_levels.  boxes: CUDA float32 (R,4) or (R,5)."""
    _require_cuda(boxes, "rois")
    if boxes.dim() != 2 or boxes.shape[1] not in (4, 5):
        raise TypeError("rois must have shape (R,4) or (R,5)")
    boxes = boxes.contiguous()
    n = boxes.shape[0]
    thr = level_thresholds(k_min, k_max)
    arr = (ctypes.c_float * max(len(thr), 1))(*thr)
    out = torch.empty(n, dtype=torch.int32 if as_int else torch.float32, device=boxes.device)
    k_cap = k_max if k_cap is None else min(k_cap, k_max)
    with torch.cuda.device(boxes.device):
        _lib.check(_lib.lib().rpool_assign_levels(
            boxes.data_ptr(), n, boxes.shape[1], roi_format, arr, len(thr), k_min, k_cap,
            None if as_int else out.data_ptr(), out.data_ptr() if as_int else None, _stream()))
    return out
