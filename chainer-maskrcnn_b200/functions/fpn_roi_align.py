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


def _image_groups(indices_and_rois, n_images, max_groups=4):
    """When the RoIs arrive image-major (batch indices non-decreasing, the order the
    reference's heads produce), returns [(n0, n1, r0, r1)]: contiguous image ranges and the
    RoI row range of each -- the call is then pipelined group by group.  One group otherwise."""
    import numpy as np
    R = indices_and_rois.shape[0]
    whole = [(0, n_images, 0, R)]
    if n_images < 2 or R == 0 or max_groups < 2:
        return whole
    img = indices_and_rois[:, 0]
    if not (np.all(img[1:] >= img[:-1]) and img[0] >= 0 and img[-1] < n_images
            and np.all(img == np.floor(img))):
        return whole
    bounds = np.linspace(0, n_images, min(n_images, max_groups) + 1).round().astype(np.int64)
    rows = np.searchsorted(img, bounds.astype(img.dtype), side="left")
    rows[-1] = R
    return [(int(bounds[i]), int(bounds[i + 1]), int(rows[i]), int(rows[i + 1]))
            for i in range(len(bounds) - 1) if bounds[i + 1] > bounds[i]]


def fpn_roi_align_host(x, indices_and_rois, levels, spatial_scales, out_sizes,
                       sampling_ratio=1, gys=None, max_groups=4):
    """Same call on NumPy host arrays (NCHW float32, as the reference holds them):
    every array is copied to the GPU, pooled there and copied back.  With ``gys``
    (one upstream gradient per pooled size) the backward pass runs too.

    The phases overlap on the two copy engines: the pooled maps travel back while the
    upstream gradients travel in, and the kernels run on the caller's stream in
    between.  When the RoIs are image-major the call is also pipelined over groups of
    images (the path shards by image: a group needs only its own pyramids, RoIs and
    gradients), so the first results leave while later inputs still arrive and only
    the last group's feature gradients are left when the uploads end.  Page-locked
    inputs are copied from where they lie; pageable ones go through a pinned bounce
    buffer.

    Returns (pooled_list, grads_list_or_None) of NumPy arrays with the
    reference's logical shapes (channels-last strides, pinned memory)."""
    import numpy as np
    single = not isinstance(out_sizes, list)
    sizes = _engine._norm_sizes([out_sizes] if single else out_sizes)
    dev = _host.device()
    cur = torch.cuda.current_stream(dev)
    s_in, s_out = _host.copy_streams(dev)
    x = [np.asarray(f) for f in x]
    rois_h = np.ascontiguousarray(indices_and_rois, dtype=np.float32)
    N, C, R = x[0].shape[0], x[0].shape[1], rois_h.shape[0]
    groups = _image_groups(rois_h, N, max_groups)

    def keep(t, *streams):          # tensors cross streams: tell the caching allocator
        for st in streams:
            t.record_stream(st)
        return t

    # results land in one pinned buffer per output; a group owns a contiguous slice of it
    pooled_h = [torch.empty_strided((R, C, oh, ow), (oh * ow * C, 1, ow * C, C), dtype=torch.float32,
                                    pin_memory=True) for oh, ow in sizes]
    grads_h = None
    if gys is not None:
        grads_h = [torch.empty_strided(f.shape, (f.shape[1] * f.shape[2] * f.shape[3], 1,
                                                 f.shape[3] * f.shape[1], f.shape[1]),
                                       dtype=torch.float32, pin_memory=True) for f in x]

    s_in.wait_stream(cur)
    staged = []
    with torch.cuda.stream(s_in):
        for n0, n1, r0, r1 in groups:
            feats = [keep(_host.h2d(f[n0:n1], dev=dev), cur) for f in x]
            r = rois_h[r0:r1]
            if n0:
                r = r.copy()
                r[:, 0] -= n0
            rois = keep(_host.h2d(r, dev=dev), cur)
            lv = None
            if levels is not None:
                lh = np.asarray(levels)[r0:r1]
                lv = _host.h2d(lh, dtype=lh.dtype if lh.dtype.kind == "i" else "float32", dev=dev)
                if lv.dtype not in (torch.int32, torch.float32):
                    lv = lv.to(torch.int32)
                keep(lv, cur)
            staged.append([feats, rois, lv, s_in.record_event(), None, None])
        if gys is not None:
            for st, (n0, n1, r0, r1) in zip(staged, groups):
                st[4] = [keep(_host.h2d(np.asarray(g)[r0:r1], dev=dev), cur) for g in gys]
                st[5] = s_in.record_event()
    plans = []
    for (feats, rois, lv, up1, _, _), (n0, n1, r0, r1) in zip(staged, groups):
        cur.wait_event(up1)
        outs, plan = _engine.forward(feats, rois, lv, list(spatial_scales)[:len(x)], sizes,
                                     sampling_ratio=sampling_ratio, roi_format=_lib.ROI_YX)
        plans.append(plan)
        s_out.wait_event(cur.record_event())
        with torch.cuda.stream(s_out):
            for o, ph in zip(outs, pooled_h):
                ph[r0:r1].copy_(keep(o, s_out), non_blocking=True)
    if gys is not None:
        for plan, (_, _, _, _, g_dev, up2), (n0, n1, r0, r1) in zip(plans, staged, groups):
            cur.wait_event(up2)
            g = _engine.backward(plan, g_dev)
            s_out.wait_event(cur.record_event())
            with torch.cuda.stream(s_out):
                for t, gh in zip(g, grads_h):
                    gh[n0:n1].copy_(keep(t, s_out), non_blocking=True)
    s_out.synchronize()
    pooled = [t.numpy() for t in pooled_h]
    grads = [t.numpy() for t in grads_h] if grads_h is not None else None
    return pooled, grads
