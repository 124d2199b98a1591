"""Drop-in for map_rois_to_fpn_levels
(chainer_maskrcnn/model/rpn/multilevel_region_proposal_network.py:16-31).

Only the level mapper of that module is on the RoI-pooling path; the proposal
network itself (convolutions, NMS) is out of scope.

The reference evaluates floor(4 + log2(sqrt(area)/224 + 1e-6)) in float32 and
clips to [k_min, k_max].  A device log2f is not guaranteed to round like
NumPy's, and one flipped floor() changes the pyramid level, so the device code
compares the float32 area against thresholds tabulated from the NumPy
expression itself: bit-exact by construction, and no device->host sync is
needed before the heads (fpn_roi_mask_head.py:58 does to_cpu on the levels).
"""
from ... import _engine, _host, _lib


def map_rois_to_fpn_levels(rois, k_min=0, k_max=4):
    """Determine which FPN level each RoI maps to (heuristic of the FPN paper).
    rois: (R, 4) y_min, x_min, y_max, x_max -- CUDA float32 tensor, or a NumPy array
    (staged through the GPU).  Returns float32 levels like the reference."""
    if _host.is_host_array(rois):
        out = _engine.assign_levels(_host.h2d(rois), _lib.ROI_YX, k_min, k_max)
        return _host.d2h(out)
    return _engine.assign_levels(rois, _lib.ROI_YX, k_min, k_max)


# ---------------------------------------------------------------------------
# cfg 5: the proposal network around the level mapper, restated on
# torch (dense cuDNN work + torchvision NMS; not part of librpool_b200.so).
# ---------------------------------------------------------------------------
import math  # noqa: E402

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def levels_by_formula(rois, k_min=0, k_max=4):
    """The reference expression in torch float32 -- for the comparison pooling
    back ends only (it may differ from NumPy in the last ulp of log2; the device
    mapper above is the bit-exact one)."""
    area = (rois[:, 2] - rois[:, 0]) * (rois[:, 3] - rois[:, 1])
    k = torch.floor(4 + torch.log2(torch.sqrt(area) / 224 + 1e-6))
    return torch.clamp(k, k_min, k_max).to(torch.float32)


def generate_anchor_base(base_size=16, ratios=(0.5, 1, 2), anchor_scales=(8,)):
    """chainercv generate_anchor_base: (len(ratios)*len(scales), 4) y1,x1,y2,x2
    boxes centred on (base_size/2, base_size/2)."""
    c = base_size / 2.
    out = []
    for r in ratios:
        for s in anchor_scales:
            h = base_size * s * math.sqrt(r)
            w = base_size * s * math.sqrt(1. / r)
            out.append([c - h / 2., c - w / 2., c + h / 2., c + w / 2.])
    return torch.tensor(out, dtype=torch.float32)


def enumerate_shifted_anchor(anchor_base, feat_stride, height, width):
    sy = torch.arange(0, height * feat_stride, feat_stride, dtype=torch.float32, device=anchor_base.device)
    sx = torch.arange(0, width * feat_stride, feat_stride, dtype=torch.float32, device=anchor_base.device)
    gy, gx = torch.meshgrid(sy, sx, indexing="ij")
    shift = torch.stack((gy.reshape(-1), gx.reshape(-1), gy.reshape(-1), gx.reshape(-1)), dim=1)
    return (shift[:, None, :] + anchor_base[None, :, :]).reshape(-1, 4)


def loc2bbox(src, loc):
    h = src[:, 2] - src[:, 0]
    w = src[:, 3] - src[:, 1]
    cy = src[:, 0] + 0.5 * h
    cx = src[:, 1] + 0.5 * w
    ncy = loc[:, 0] * h + cy
    ncx = loc[:, 1] * w + cx
    nh = torch.exp(loc[:, 2].clamp(max=10.)) * h
    nw = torch.exp(loc[:, 3].clamp(max=10.)) * w
    return torch.stack((ncy - 0.5 * nh, ncx - 0.5 * nw, ncy + 0.5 * nh, ncx + 0.5 * nw), dim=1)


def bbox2loc(src, dst):
    eps = torch.finfo(torch.float32).eps
    h = (src[:, 2] - src[:, 0]).clamp(min=eps)
    w = (src[:, 3] - src[:, 1]).clamp(min=eps)
    cy = src[:, 0] + 0.5 * (src[:, 2] - src[:, 0])
    cx = src[:, 1] + 0.5 * (src[:, 3] - src[:, 1])
    bh = dst[:, 2] - dst[:, 0]
    bw = dst[:, 3] - dst[:, 1]
    bcy = dst[:, 0] + 0.5 * bh
    bcx = dst[:, 1] + 0.5 * bw
    return torch.stack(((bcy - cy) / h, (bcx - cx) / w, torch.log(bh / h), torch.log(bw / w)), dim=1)


class ProposalCreator(object):
    """chainer_maskrcnn/utils/proposal_creator.py:53-169 on the device: decode,
    clip, size filter, top-k, NMS (torchvision), top-k."""

    def __init__(self, nms_thresh=0.7, n_train_pre_nms=12000, n_train_post_nms=2000,
                 n_test_pre_nms=6000, n_test_post_nms=300, min_size=16):
        self.nms_thresh = nms_thresh
        self.n_train_pre_nms, self.n_train_post_nms = n_train_pre_nms, n_train_post_nms
        self.n_test_pre_nms, self.n_test_post_nms = n_test_pre_nms, n_test_post_nms
        self.min_size = min_size

    @torch.no_grad()
    def __call__(self, loc, score, anchor, img_size, scale=1., train=True):
        from torchvision.ops import nms
        n_pre = self.n_train_pre_nms if train else self.n_test_pre_nms
        n_post = self.n_train_post_nms if train else self.n_test_post_nms
        roi = loc2bbox(anchor, loc)
        roi[:, 0::2] = roi[:, 0::2].clamp(0, img_size[0])
        roi[:, 1::2] = roi[:, 1::2].clamp(0, img_size[1])
        min_size = self.min_size * scale
        keep = ((roi[:, 2] - roi[:, 0]) >= min_size) & ((roi[:, 3] - roi[:, 1]) >= min_size)
        roi, score = roi[keep], score[keep]
        if n_pre > 0 and score.numel() > n_pre:
            score, order = torch.topk(score, n_pre)
        else:
            score, order = torch.sort(score, descending=True)
        roi = roi[order]
        keep = nms(roi[:, [1, 0, 3, 2]], score, self.nms_thresh)
        if n_post > 0:
            keep = keep[:n_post]
        return roi[keep]


class MultilevelRegionProposalNetwork(nn.Module):
    """chainer_maskrcnn/model/rpn/multilevel_region_proposal_network.py:34-166:
    one 3x3 conv + score/loc 1x1 convs shared by every level, one anchor scale per
    level, proposals per image, then the level of every proposal."""

    def __init__(self, anchor_scales, feat_strides, in_channels=256, mid_channels=256,
                 ratios=(0.5, 1, 2), proposal_creator_params=None, level_fn=None):
        super().__init__()
        if len(anchor_scales) != len(feat_strides):
            raise ValueError('length of anchor_scales and feat_strides should be same!')
        self.anchor_bases = [generate_anchor_base(anchor_scales=[s], ratios=ratios) for s in anchor_scales]
        self.feat_strides = list(feat_strides)
        self.proposal_layer = ProposalCreator(**(proposal_creator_params or {}))
        n_anchor = self.anchor_bases[0].shape[0]
        self.conv = nn.Conv2d(in_channels, mid_channels, 3, 1, 1)
        self.score = nn.Conv2d(mid_channels, n_anchor * 2, 1, 1, 0)
        self.loc = nn.Conv2d(mid_channels, n_anchor * 4, 1, 1, 0)
        for m in (self.conv, self.score, self.loc):
            nn.init.normal_(m.weight, std=0.01)
            nn.init.zeros_(m.bias)
        # None: the device mapper of this module (bit-exact); tests on the CPU pass
        # levels_by_formula
        self.level_fn = level_fn

    def forward(self, xs, img_size, scale=1.):
        locs, scores, fg_scores, anchors = [], [], [], []
        for i, x in enumerate(xs):
            n, _, hh, ww = x.shape
            anchor = enumerate_shifted_anchor(self.anchor_bases[i].to(x.device), self.feat_strides[i], hh, ww)
            n_anchor = anchor.shape[0] // (hh * ww)
            h = F.relu(self.conv(x))
            rpn_locs = self.loc(h).permute(0, 2, 3, 1).reshape(n, -1, 4)
            rpn_scores = self.score(h).permute(0, 2, 3, 1)
            rpn_fg = rpn_scores.reshape(n, hh, ww, n_anchor, 2)[..., 1].reshape(n, -1)
            locs.append(rpn_locs)
            scores.append(rpn_scores.reshape(n, -1, 2))
            fg_scores.append(rpn_fg)
            anchors.append(anchor)
        locs = torch.cat(locs, dim=1)
        scores = torch.cat(scores, dim=1)
        fg_scores = torch.cat(fg_scores, dim=1)
        anchors = torch.cat(anchors, dim=0)
        rois, roi_indices = [], []
        for i in range(locs.shape[0]):
            roi = self.proposal_layer(locs[i].detach().float(), fg_scores[i].detach().float(), anchors,
                                      img_size, scale=scale, train=self.training)
            rois.append(roi)
            roi_indices.append(torch.full((roi.shape[0],), i, dtype=torch.int32, device=roi.device))
        rois = torch.cat(rois, dim=0)
        roi_indices = torch.cat(roi_indices, dim=0)
        levels = (self.level_fn or map_rois_to_fpn_levels)(rois)
        return locs, scores, rois, roi_indices, anchors, levels
