"""FPNRoIMaskHead on torch (cfg 5): the layers of
chainer_maskrcnn/model/head/fpn_roi_mask_head.py:13-102 around the pooling path.

The conv / fc / deconv layers are dense cuDNN/cuBLAS work and are not part of
librpool_b200.so; what this module is for is the ``pooling`` switch, which selects
how the two pooled tensors of the training call (:57-63 box 7x7, :74-78 mask
14x14) are produced:

    "b200"         one fused launch of this package's kernels for both sizes
                   (FPNRoIPooling -> fpn_roi_align): the product path;
    "torchvision"  torchvision.ops.roi_align, one batched call per level and size
                   with the rows scattered back into input order (how torchvision's
                   own MultiScaleRoIAlign dispatches): library comparison arm;
    "per_roi"      the reference's dispatch -- one operator call per RoI and size,
                   concatenated in input order (:57-63) -- with torchvision's kernel
                   as the operator: reference-dispatch comparison arm.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .fpn_roi_pooling import FPNRoIPooling


def _pool_torchvision(x, indices_and_rois, levels, spatial_scales, size, sampling_ratio, per_roi):
    from torchvision.ops import roi_align
    rois_xy = indices_and_rois[:, [0, 2, 1, 4, 3]]            # roi_align_2d_yx.py:4-7
    lv = levels.to(torch.int64).clamp(0, len(x) - 1)
    sr = sampling_ratio if sampling_ratio > 0 else -1
    if per_roi:
        lv_host = lv.cpu().tolist()                             # fpn_roi_mask_head.py:58
        pool = [roi_align(x[l], rois_xy[i:i + 1], size, spatial_scales[l], sr, aligned=False)
                for i, l in enumerate(lv_host)]
        return torch.cat(pool, dim=0)
    out = x[0].new_zeros((rois_xy.shape[0], x[0].shape[1], size, size))
    for l in range(len(x)):
        idx = torch.nonzero(lv == l).squeeze(1)
        if idx.numel():
            out = out.index_copy(0, idx, roi_align(x[l], rois_xy[idx], size, spatial_scales[l], sr,
                                                   aligned=False))
    return out


class FPNRoIMaskHead(nn.Module):
    mask_size = 28

    def __init__(self, n_class, roi_size_box=7, roi_size_mask=14, channels=256, fc_dim=1024,
                 pooling="b200", sampling_ratio=1):
        super().__init__()
        c = channels
        self.conv1 = nn.Conv2d(c, c, 3, 1, 1)
        self.fc1 = nn.Linear(c * roi_size_box * roi_size_box, fc_dim)
        self.fc2 = nn.Linear(fc_dim, fc_dim)
        self.cls_loc = nn.Linear(fc_dim, 4)
        self.score = nn.Linear(fc_dim, n_class)
        self.mask1 = nn.Conv2d(c, c, 3, 1, 1)
        self.mask2 = nn.Conv2d(c, c, 3, 1, 1)
        self.mask3 = nn.Conv2d(c, c, 3, 1, 1)
        self.mask4 = nn.Conv2d(c, c, 3, 1, 1)
        self.deconv1 = nn.ConvTranspose2d(c, c, 2, 2, 0)
        self.conv2 = nn.Conv2d(c, n_class - 1, 1, 1, 0)
        nn.init.normal_(self.cls_loc.weight, std=0.001)
        nn.init.normal_(self.score.weight, std=0.01)
        for m in (self.deconv1, self.conv2):
            nn.init.normal_(m.weight, std=0.01)
        self.n_class = n_class
        self.roi_size_box = roi_size_box
        self.roi_size_mask = roi_size_mask
        self.sampling_ratio = sampling_ratio
        if pooling not in ("b200", "torchvision", "per_roi"):
            raise ValueError("unknown pooling back end %r" % (pooling,))
        self.pooling = pooling
        self.pool = FPNRoIPooling(roi_size_box, roi_size_mask, sampling_ratio)
        self.x = None

    def _pool(self, x, indices_and_rois, levels, spatial_scales, size):
        return _pool_torchvision(x, indices_and_rois, levels, spatial_scales, size, self.sampling_ratio,
                                 per_roi=(self.pooling == "per_roi"))

    def _mask_branch(self, pool_mask):
        mask = F.relu(self.mask1(pool_mask))
        mask = F.relu(self.mask2(mask))
        mask = F.relu(self.mask3(mask))
        mask = F.relu(self.mask4(mask))
        return self.conv2(self.deconv1(mask))

    def forward(self, x, indices_and_rois, levels, spatial_scales):
        x = [f.float() for f in x]
        pool_mask = None
        if self.pooling == "b200":
            if self.training:
                pool_box, pool_mask = self.pool(x, indices_and_rois, levels, spatial_scales, train=True)
            else:
                pool_box = self.pool(x, indices_and_rois, levels, spatial_scales, train=False)
        else:
            pool_box = self._pool(x, indices_and_rois, levels, spatial_scales, self.roi_size_box)
            if self.training:
                pool_mask = self._pool(x, indices_and_rois, levels, spatial_scales, self.roi_size_mask)
        h = F.relu(self.conv1(pool_box))
        h = F.relu(self.fc1(h.flatten(1)))
        h = F.relu(self.fc2(h))
        roi_cls_locs = self.cls_loc(h)
        roi_scores = self.score(h)
        if self.training:
            return roi_cls_locs, roi_scores, self._mask_branch(pool_mask)
        self.x = x   # cache for the second pass (:85-87)
        return roi_cls_locs, roi_scores

    def predict_mask(self, levels, indices_and_rois, spatial_scales):
        if self.pooling == "b200":
            pool_mask = self.pool.predict_mask(levels, indices_and_rois, spatial_scales)
        else:
            pool_mask = self._pool(self.x, indices_and_rois, levels, spatial_scales, self.roi_size_mask)
        return self._mask_branch(pool_mask)
