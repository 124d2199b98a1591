"""FPNMaskRCNNTrainChain on torch (cfg 5): one training step of the
reference model (chainer_maskrcnn/model/fpn_maskrcnn_train_chain.py:14-117) around the
pooling path, so that the path's effect can be measured in context.

Everything here except the head's pooling call is dense library work (cuDNN, cuBLAS,
torchvision NMS) or small target bookkeeping; none of it is in librpool_b200.so.
Differences from the reference, all outside the pooling path:
  * batches of n >= 1 images (the reference raises for n != 1, :37-40): targets are
    built per image, the heads run once over all images' sampled RoIs;
  * targets are built on the device (the reference round-trips through NumPy);
  * mask targets are cut with a bilinear crop-and-resize on the device instead of
    int() cropping + cv2.resize (proposal_target_creator.py:97-108).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .extractor.feature_pyramid_network import FeaturePyramidNetwork
from .head.fpn_roi_mask_head import FPNRoIMaskHead
from .rpn.multilevel_region_proposal_network import (MultilevelRegionProposalNetwork, bbox2loc,
                                                     levels_by_formula, map_rois_to_fpn_levels)


def bbox_iou(a, b):
    tl = torch.maximum(a[:, None, :2], b[None, :, :2])
    br = torch.minimum(a[:, None, 2:], b[None, :, 2:])
    inter = (br - tl).clamp(min=0).prod(dim=2) * (tl < br).all(dim=2)
    area_a = (a[:, 2:] - a[:, :2]).prod(dim=1)
    area_b = (b[:, 2:] - b[:, :2]).prod(dim=1)
    return inter / (area_a[:, None] + area_b[None, :] - inter)


def _choice(index, n, gen):
    """np.random.choice(index, size=n, replace=False)"""
    if index.numel() <= n:
        return index
    perm = torch.randperm(index.numel(), device=index.device, generator=gen)[:n]
    return index[perm]


class AnchorTargetCreator(object):
    """chainercv AnchorTargetCreator (used at fpn_maskrcnn_train_chain.py:80-82)."""

    def __init__(self, n_sample=256, pos_iou_thresh=0.7, neg_iou_thresh=0.3, pos_ratio=0.5):
        self.n_sample, self.pos_ratio = n_sample, pos_ratio
        self.pos_iou_thresh, self.neg_iou_thresh = pos_iou_thresh, neg_iou_thresh

    @torch.no_grad()
    def __call__(self, bbox, anchor, img_size, gen=None):
        H, W = img_size
        n_anchor = anchor.shape[0]
        inside = torch.nonzero((anchor[:, 0] >= 0) & (anchor[:, 1] >= 0) &
                               (anchor[:, 2] <= H) & (anchor[:, 3] <= W)).squeeze(1)
        a = anchor[inside]
        iou = bbox_iou(a, bbox)
        max_iou, argmax = iou.max(dim=1)
        gt_max = iou.max(dim=0).values
        label = torch.full((a.shape[0],), -1, dtype=torch.int64, device=a.device)
        label[max_iou < self.neg_iou_thresh] = 0
        label[(iou == gt_max[None, :]).any(dim=1)] = 1
        label[max_iou >= self.pos_iou_thresh] = 1
        n_pos = int(self.pos_ratio * self.n_sample)
        pos = torch.nonzero(label == 1).squeeze(1)
        if pos.numel() > n_pos:
            drop = pos[torch.randperm(pos.numel(), device=a.device, generator=gen)[:pos.numel() - n_pos]]
            label[drop] = -1
        n_neg = self.n_sample - int((label == 1).sum())
        neg = torch.nonzero(label == 0).squeeze(1)
        if neg.numel() > n_neg:
            drop = neg[torch.randperm(neg.numel(), device=a.device, generator=gen)[:neg.numel() - n_neg]]
            label[drop] = -1
        loc = bbox2loc(a, bbox[argmax])
        full_label = torch.full((n_anchor,), -1, dtype=torch.int64, device=a.device)
        full_label[inside] = label
        full_loc = torch.zeros((n_anchor, 4), dtype=torch.float32, device=a.device)
        full_loc[inside] = loc
        return full_loc, full_label


class ProposalTargetCreator(object):
    """chainer_maskrcnn/utils/proposal_target_creator.py:11-137 (binary-mask branch)."""

    def __init__(self, n_sample=256, pos_ratio=0.25, pos_iou_thresh=0.5,
                 neg_iou_thresh_hi=0.5, neg_iou_thresh_lo=0.0, level_fn=None):
        self.n_sample, self.pos_ratio = n_sample, pos_ratio
        self.pos_iou_thresh = pos_iou_thresh
        self.neg_iou_thresh_hi, self.neg_iou_thresh_lo = neg_iou_thresh_hi, neg_iou_thresh_lo
        self.level_fn = level_fn

    @torch.no_grad()
    def __call__(self, roi, bbox, label, mask, levels, loc_normalize_mean=(0., 0., 0., 0.),
                 loc_normalize_std=(0.1, 0.1, 0.2, 0.2), mask_size=14, gen=None):
        from torchvision.ops import roi_align as _crop_resize
        roi = torch.cat((roi, bbox), dim=0)                                   # :46
        bbox_levels = (self.level_fn or map_rois_to_fpn_levels)(bbox)         # :49
        levels = torch.cat((levels, bbox_levels))                             # :50
        pos_roi_per_image = round(self.n_sample * self.pos_ratio)
        iou = bbox_iou(roi, bbox)
        max_iou, gt_assignment = iou.max(dim=1)
        gt_roi_label = label[gt_assignment] + 1
        pos_index = torch.nonzero(max_iou >= self.pos_iou_thresh).squeeze(1)
        n_pos = int(min(pos_roi_per_image, pos_index.numel()))
        pos_index = _choice(pos_index, n_pos, gen)
        neg_index = torch.nonzero((max_iou < self.neg_iou_thresh_hi) &
                                  (max_iou >= self.neg_iou_thresh_lo)).squeeze(1)
        n_neg = int(min(self.n_sample - n_pos, neg_index.numel()))
        neg_index = _choice(neg_index, n_neg, gen)
        keep_index = torch.cat((pos_index, neg_index))
        gt_roi_label = gt_roi_label[keep_index].clone()
        gt_roi_label[n_pos:] = 0
        sample_roi = roi[keep_index]
        sample_levels = levels[keep_index]
        gt_roi_loc = bbox2loc(sample_roi, bbox[gt_assignment[keep_index]])
        mean = torch.tensor(loc_normalize_mean, dtype=torch.float32, device=roi.device)
        std = torch.tensor(loc_normalize_std, dtype=torch.float32, device=roi.device)
        gt_roi_loc = (gt_roi_loc - mean) / std
        # mask targets of the positives: their ground-truth mask cut to the RoI and
        # resized to mask_size x mask_size
        if n_pos:
            pr = sample_roi[:n_pos]
            boxes = torch.stack((gt_assignment[pos_index].float(), pr[:, 1], pr[:, 0], pr[:, 3], pr[:, 2]), 1)
            crop = _crop_resize(mask[:, None].float(), boxes, mask_size, 1.0, 2, aligned=False)
            # the image index column selects the instance's own mask plane
            gt_roi_mask = (crop[:, 0] >= 0.5).to(torch.int64)
        else:
            gt_roi_mask = torch.zeros((0, mask_size, mask_size), dtype=torch.int64, device=roi.device)
        return sample_roi, sample_levels, gt_roi_loc, gt_roi_label, gt_roi_mask


def _smooth_l1_loss(x, t, in_weight, sigma):
    sigma2 = sigma ** 2
    diff = in_weight * (x - t)
    abs_diff = diff.abs()
    flag = (abs_diff < (1. / sigma2)).float()
    y = flag * (sigma2 / 2.) * diff * diff + (1 - flag) * (abs_diff - 0.5 / sigma2)
    return y.sum()


def _fast_rcnn_loc_loss(pred_loc, gt_loc, gt_label, sigma):
    in_weight = torch.zeros_like(gt_loc)
    in_weight[gt_label > 0] = 1
    loc_loss = _smooth_l1_loss(pred_loc, gt_loc, in_weight, sigma)
    return loc_loss / (gt_label >= 0).sum().clamp(min=1)


def calc_mask_loss(roi_cls_mask, gt_roi_mask, gt_roi_label):
    """train.py:50-58: sigmoid cross entropy of the ground-truth class' mask plane."""
    n = gt_roi_mask.shape[0]
    if n == 0:
        return roi_cls_mask.sum() * 0
    idx = torch.arange(n, device=roi_cls_mask.device)
    roi_mask = roi_cls_mask[idx, gt_roi_label[:n] - 1]
    return F.binary_cross_entropy_with_logits(roi_mask, gt_roi_mask.float())


class MaskRCNN(nn.Module):
    """extractor + rpn + head as chainer_maskrcnn/model/maskrcnn.py:53-58,98-105 wires
    them for backbone='fpn', head_arch='fpn'."""

    def __init__(self, n_fg_class, pooling="b200", sampling_ratio=1, width=64, blocks=(3, 4, 6, 3),
                 channels=256, fc_dim=1024, proposal_creator_params=None, level_fn=None):
        super().__init__()
        self.extractor = FeaturePyramidNetwork(width, blocks, channels)
        self.rpn = MultilevelRegionProposalNetwork(
            self.extractor.anchor_scales, self.extractor.feat_strides, channels, channels,
            proposal_creator_params=proposal_creator_params, level_fn=level_fn)
        self.head = FPNRoIMaskHead(n_fg_class + 1, 7, 14, channels, fc_dim, pooling, sampling_ratio)
        self.n_class = n_fg_class + 1
        self.loc_normalize_mean = (0., 0., 0., 0.)
        self.loc_normalize_std = (0.1, 0.1, 0.2, 0.2)


class FPNMaskRCNNTrainChain(nn.Module):
    def __init__(self, faster_rcnn, rpn_sigma=3., roi_sigma=1., level_fn=None, seed=0):
        super().__init__()
        self.faster_rcnn = faster_rcnn
        self.rpn_sigma, self.roi_sigma = rpn_sigma, roi_sigma
        self.anchor_target_creator = AnchorTargetCreator()
        self.proposal_target_creator = ProposalTargetCreator(level_fn=level_fn)
        self.seed = seed
        self._gen = None
        self.last = {}

    def _generator(self, device):
        if self._gen is None or self._gen.device != device:
            self._gen = torch.Generator(device=device)
            self._gen.manual_seed(self.seed)
        return self._gen

    def forward(self, imgs, bboxes, labels, masks, scale=1.):
        """imgs (n,3,H,W); bboxes/labels/masks: per-image lists of (G,4) y1,x1,y2,x2,
        (G,) int64 in [0, n_fg_class), (G,H,W) bool."""
        m = self.faster_rcnn
        n, _, H, W = imgs.shape
        img_size = (H, W)
        gen = self._generator(imgs.device)
        features = m.extractor(imgs)
        rpn_locs, rpn_scores, rois, roi_indices, anchor, levels = m.rpn(features, img_size, scale)

        samples, s_levels, gt_locs, gt_labels, gt_masks, pos_rows = [], [], [], [], [], []
        rpn_loc_loss = rpn_cls_loss = 0.
        row0 = 0
        for i in range(n):
            sel = roi_indices == i
            sample_roi, sample_levels, gt_roi_loc, gt_roi_label, gt_roi_mask = self.proposal_target_creator(
                rois[sel], bboxes[i], labels[i], masks[i], levels[sel], m.loc_normalize_mean,
                m.loc_normalize_std, mask_size=m.head.mask_size, gen=gen)
            idx = torch.full((sample_roi.shape[0], 1), float(i), device=imgs.device)
            samples.append(torch.cat((idx, sample_roi), dim=1).float())       # :73-78
            s_levels.append(sample_levels)
            gt_locs.append(gt_roi_loc)
            gt_labels.append(gt_roi_label)
            gt_masks.append(gt_roi_mask)
            pos_rows.append(row0 + torch.arange(gt_roi_mask.shape[0], device=imgs.device))
            row0 += sample_roi.shape[0]
            gt_rpn_loc, gt_rpn_label = self.anchor_target_creator(bboxes[i], anchor, img_size, gen=gen)
            rpn_loc_loss = rpn_loc_loss + _fast_rcnn_loc_loss(rpn_locs[i].float(), gt_rpn_loc, gt_rpn_label,
                                                              self.rpn_sigma) / n
            rpn_cls_loss = rpn_cls_loss + F.cross_entropy(rpn_scores[i].float(), gt_rpn_label,
                                                          ignore_index=-1) / n
        indices_and_rois = torch.cat(samples, dim=0)
        sample_levels = torch.cat(s_levels)
        gt_roi_loc = torch.cat(gt_locs)
        gt_roi_label = torch.cat(gt_labels)
        gt_roi_mask = torch.cat(gt_masks)
        pos_rows = torch.cat(pos_rows)

        roi_cls_loc, roi_score, roi_cls_mask = m.head(
            features, indices_and_rois, sample_levels, m.extractor.spatial_scales)
        n_sample = roi_cls_loc.shape[0]
        roi_cls_loc = roi_cls_loc.float().reshape(n_sample, -1, 4)
        if roi_cls_loc.shape[1] == 1:
            roi_loc = roi_cls_loc.reshape(n_sample, 4)
        else:
            roi_loc = roi_cls_loc[torch.arange(n_sample, device=imgs.device), gt_roi_label]
        roi_loc_loss = _fast_rcnn_loc_loss(roi_loc, gt_roi_loc, gt_roi_label, self.roi_sigma)
        roi_cls_loss = F.cross_entropy(roi_score.float(), gt_roi_label)
        mask_loss = calc_mask_loss(roi_cls_mask.float()[pos_rows], gt_roi_mask, gt_roi_label[pos_rows])
        loss = rpn_loc_loss + rpn_cls_loss + roi_loc_loss + roi_cls_loss + mask_loss
        self.last = {"rpn_loc_loss": rpn_loc_loss, "rpn_cls_loss": rpn_cls_loss,
                     "roi_loc_loss": roi_loc_loss, "roi_cls_loss": roi_cls_loss, "mask_loss": mask_loss,
                     "loss": loss, "n_sample": n_sample, "n_proposals": int(rois.shape[0]),
                     "features": features, "indices_and_rois": indices_and_rois, "levels": sample_levels}
        return loss


def synthetic_batch(n_images, height, width, n_fg_class=80, max_gt=8, seed=0, device="cpu"):
    """COCO-shaped synthetic batch: images ~N(0,1); 2..max_gt boxes per image with
    sqrt(area) log-uniform in [32, 400] px; masks = the ellipse inscribed in each box."""
    g = torch.Generator().manual_seed(seed)
    imgs = torch.randn((n_images, 3, height, width), generator=g)
    bboxes, labels, masks = [], [], []
    yy = torch.arange(height, dtype=torch.float32)[:, None]
    xx = torch.arange(width, dtype=torch.float32)[None, :]
    for _ in range(n_images):
        G = int(torch.randint(2, max_gt + 1, (1,), generator=g))
        s = torch.exp(torch.rand(G, generator=g) * (torch.log(torch.tensor(400.)) - torch.log(torch.tensor(32.)))
                      + torch.log(torch.tensor(32.)))
        ar = torch.exp((torch.rand(G, generator=g) - 0.5) * 2 * torch.log(torch.tensor(2.)))
        h = (s * torch.sqrt(ar)).clamp(max=height - 1.)
        w = (s / torch.sqrt(ar)).clamp(max=width - 1.)
        y1 = torch.rand(G, generator=g) * (height - h)
        x1 = torch.rand(G, generator=g) * (width - w)
        box = torch.stack((y1, x1, y1 + h, x1 + w), dim=1)
        cy, cx = (y1 + h / 2)[:, None, None], (x1 + w / 2)[:, None, None]
        msk = (((yy[None] - cy) / (h[:, None, None] / 2)) ** 2 + ((xx[None] - cx) / (w[:, None, None] / 2)) ** 2) <= 1
        bboxes.append(box.to(device))
        labels.append(torch.randint(0, n_fg_class, (G,), generator=g).to(device))
        masks.append(msk.to(device))
    return imgs.to(device), bboxes, labels, masks
