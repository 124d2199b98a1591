from .fpn_roi_pooling import FPNRoIPooling  # noqa: F401
