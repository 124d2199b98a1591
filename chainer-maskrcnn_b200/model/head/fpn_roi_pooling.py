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


class FPNRoIKeypointPooling(FPNRoIPooling):
    """Pooling lines of FPNRoIKeypointHead (fpn_roi_keypoint_head.py:57-111).

    They differ from the mask head in one place: when every RoI carries the same
    level, the box features are pooled with ONE batched call on ``x[0]`` at
    ``spatial_scales[0]`` -- whatever that common level is (:62-64; right for
    single-level backbones only).  The mask branch always follows the per-RoI
    level (:83-87, :99-104).  ``reference_quirk=True`` (default) reproduces that,
    at the price of the same device->host read of ``levels`` the reference pays
    (:60); ``False`` pools box and mask by level in one launch like the mask head.
    """

    def __init__(self, roi_size_box=7, roi_size_mask=14, sampling_ratio=1, reference_quirk=True):
        super(FPNRoIKeypointPooling, self).__init__(roi_size_box, roi_size_mask, sampling_ratio)
        self.reference_quirk = reference_quirk

    def _single_level(self, levels):
        if not self.reference_quirk or levels is None or levels.numel() == 0:
            return False
        lv = levels.int()                            # .astype(np.int32), :60
        return bool((lv == lv[0]).all().item())      # len(np.unique(levels)) == 1, :62

    def __call__(self, x, indices_and_rois, levels, spatial_scales, train=True):
        if not self._single_level(levels):
            return super(FPNRoIKeypointPooling, self).__call__(x, indices_and_rois, levels,
                                                               spatial_scales, train)
        pool_box = fpn_roi_align(x[:1], indices_and_rois, None, spatial_scales[:1],
                                 self.roi_size_box, self.sampling_ratio)
        if not train:
            self.x = x
            return pool_box
        pool_mask = fpn_roi_align(x, indices_and_rois, levels, spatial_scales,
                                  self.roi_size_mask, self.sampling_ratio)
        return pool_box, pool_mask
