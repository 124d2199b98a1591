from .roi_align_2d import ROIAlign2D, roi_align_2d, InvalidType  # noqa: F401
