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
