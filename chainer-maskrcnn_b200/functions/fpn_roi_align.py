"""Fused replacement of the FPN heads' per-RoI dispatch loops.

The reference pools RoI r from pyramid level levels[r] with one operator call
per RoI and concatenates the results in input order:

    for l, i in zip(levels, indices_and_rois):
        pool.append(_roi_align_2d_yx(x[l], i[None], S, S, spatial_scales[l]))
    pool = F.concat(pool, axis=0)
        chainer_maskrcnn/model/head/fpn_roi_mask_head.py:57-63 (box, S=7),
        :74-78 and :90-95 (mask, S=14); fpn_roi_keypoint_head.py:59-71,83-87,99-104

``fpn_roi_align`` takes the same four arguments the heads receive
(x, indices_and_rois, levels, spatial_scales) plus the pooled size(s) and does
all of it -- every level, every RoI, one or two pooled sizes -- in a single
kernel launch; row r of each result is RoI r of the input.
"""
import torch

from .. import _engine, _host, _lib


def fpn_roi_align(x, indices_and_rois, levels, spatial_scales, out_sizes,
                  sampling_ratio=1, coord_mode=None):
    """x: tuple of (N,C,H_l,W_l) float32 CUDA tensors (fine -> coarse);
    indices_and_rois: (R,5) float32 [batch_index, y1, x1, y2, x2] image coordinates;
    levels: (R,) float32/int32 CUDA tensor as produced by map_rois_to_fpn_levels, or
        None to assign them on the device by the same rule (clipped to len(x)-1 as
        chainer_maskrcnn/model/maskrcnn.py:141 does);
    spatial_scales: per-level 1/stride (feature_pyramid_network.py:11);
    out_sizes: int, (h, w), or a list of up to two of those (box + mask in one launch).

    Returns one (R,C,S,S) tensor, or a tuple when several sizes were asked for.
    Differentiable with respect to every level of ``x``."""
    single = not isinstance(out_sizes, list)
    sizes = [out_sizes] if single else out_sizes
    # the reference indexes spatial_scales[l] per RoI: a list longer than the pyramid is fine
    if len(spatial_scales) < len(x):
        raise ValueError("one spatial_scale per level: %d scales for %d levels" % (len(spatial_scales), len(x)))
    spatial_scales = list(spatial_scales)[:len(x)]
    cfg = dict(spatial_scales=list(spatial_scales), out_sizes=sizes,
               sampling_ratio=sampling_ratio, roi_format=_lib.ROI_YX)
    if coord_mode is not None:
        cfg["coord_mode"] = coord_mode
    outs = _engine.apply(list(x), indices_and_rois, levels, **cfg)
    return outs[0] if single else tuple(outs)


def fpn_roi_align_host(x, indices_and_rois, levels, spatial_scales, out_sizes,
                       sampling_ratio=1, gys=None):
    """Same call on NumPy host arrays (NCHW float32, as the reference holds them):
    every array is copied to the GPU, pooled there and copied back.  With ``gys``
    (one upstream gradient per pooled size) the backward pass runs too.

    The three phases overlap on the two copy engines: the pooled maps travel back
    while the upstream gradients travel in, and the kernels run on the caller's
    stream in between.  Page-locked inputs are copied from where they lie; pageable
    ones go through a pinned bounce buffer.

    Returns (pooled_list, grads_list_or_None) of NumPy arrays with the
    reference's logical shapes (channels-last strides, pinned memory)."""
    single = not isinstance(out_sizes, list)
    sizes = [out_sizes] if single else out_sizes
    dev = _host.device()
    cur = torch.cuda.current_stream(dev)
    s_in, s_out = _host.copy_streams(dev)

    def keep(t, *streams):          # tensors cross streams: tell the caching allocator
        for st in streams:
            t.record_stream(st)
        return t

    s_in.wait_stream(cur)
    with torch.cuda.stream(s_in):
        feats = [keep(_host.h2d(f, dev=dev), cur) for f in x]
        rois = keep(_host.h2d(indices_and_rois, dev=dev), cur)
        lv = None
        if levels is not None:
            lv = _host.h2d(levels, dtype=levels.dtype if levels.dtype.kind == "i" else "float32", dev=dev)
            if lv.dtype not in (torch.int32, torch.float32):
                lv = lv.to(torch.int32)
            keep(lv, cur)
        up1 = s_in.record_event()
        g_dev, up2 = None, None
        if gys is not None:
            g_dev = [keep(_host.h2d(g, dev=dev), cur) for g in gys]
            up2 = s_in.record_event()
    cur.wait_event(up1)
    outs, plan = _engine.forward(feats, rois, lv, list(spatial_scales), sizes,
                                 sampling_ratio=sampling_ratio, roi_format=_lib.ROI_YX)
    s_out.wait_event(cur.record_event())
    with torch.cuda.stream(s_out):
        pooled = [_host.d2h_async(keep(o, s_out)) for o in outs]
    grads = None
    if gys is not None:
        cur.wait_event(up2)
        g = _engine.backward(plan, g_dev)
        s_out.wait_event(cur.record_event())
        with torch.cuda.stream(s_out):
            grads = [_host.d2h_async(keep(t, s_out)) for t in g]
    s_out.synchronize()
    pooled = [t.numpy() for t in pooled]
    if grads is not None:
        grads = [t.numpy() for t in grads]
    return pooled, grads
