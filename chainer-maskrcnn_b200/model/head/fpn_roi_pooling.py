"""The pooling half of the reference FPN heads, as one object.

FPNRoIMaskHead / FPNRoIKeypointHead (chainer_maskrcnn/model/head/
fpn_roi_mask_head.py:55-102, fpn_roi_keypoint_head.py:57-111) interleave RoI
pooling with their conv/fc layers.  The layers are out of scope (dense cuDNN
work); this class carries the pooling lines with the same call signatures so a
head can delegate to it:

    __call__(x, indices_and_rois, levels, spatial_scales)      :55  / :57
        train  -> (pool_box, pool_mask)   (:57-63 and :74-78, one launch)
        test   -> pool_box, and caches x  (:85-87)
    predict_mask(levels, indices_and_rois, spatial_scales)     :90-95 / :99-104
"""
from ...functions.fpn_roi_align import fpn_roi_align


class FPNRoIPooling(object):
    def __init__(self, roi_size_box=7, roi_size_mask=14, sampling_ratio=1):
        self.roi_size_box = roi_size_box
        self.roi_size_mask = roi_size_mask
        self.sampling_ratio = sampling_ratio
        self.x = None

    def __call__(self, x, indices_and_rois, levels, spatial_scales, train=True):
        if train:
            # box and mask windows overlap completely: pool both from one read
            pool_box, pool_mask = fpn_roi_align(
                x, indices_and_rois, levels, spatial_scales,
                [self.roi_size_box, self.roi_size_mask], self.sampling_ratio)
            return pool_box, pool_mask
        self.x = x  # cache, fpn_roi_mask_head.py:85-87
        return fpn_roi_align(x, indices_and_rois, levels, spatial_scales,
                             self.roi_size_box, self.sampling_ratio)

    def predict_mask(self, levels, indices_and_rois, spatial_scales):
        if self.x is None:
            raise RuntimeError("predict_mask needs the features cached by a test-mode call")
        return fpn_roi_align(self.x, indices_and_rois, levels, spatial_scales,
                             self.roi_size_mask, self.sampling_ratio)
