from .fpn_roi_pooling import FPNRoIPooling, FPNRoIKeypointPooling  # noqa: F401
