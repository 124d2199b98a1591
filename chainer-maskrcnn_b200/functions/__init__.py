from .roi_align.roi_align_2d import ROIAlign2D, roi_align_2d  # noqa: F401
from .roi_align_2d_yx import _roi_align_2d_yx  # noqa: F401
from .fpn_roi_align import fpn_roi_align, fpn_roi_align_host  # noqa: F401
