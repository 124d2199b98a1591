"""Pyramid constants of the reference extractor
(chainer_maskrcnn/model/extractor/feature_pyramid_network.py:9-11,71).  The
backbone convolutions are out of scope; the RoI-pooling path only needs the
strides, the scales handed to the op, and the (p2..p6) fine-to-coarse order."""
import math

feat_strides = [4, 8, 16, 32, 64]
# inverse of feat_strides: image coordinates -> feature-map coordinates
spatial_scales = list(map(lambda x: 1. / x, feat_strides))


def pyramid_shapes(n_images, channels, height, width, n_levels=5):
    """(N, C, ceil(H/s), ceil(W/s)) per level -- what the reference FPN produces
    for an H x W input (conv1 s2 p3 -> cover-all max-pool -> stride-2 stages,
    feature_pyramid_network.py:48-53)."""
    return [(n_images, channels, math.ceil(height / s), math.ceil(width / s))
            for s in feat_strides[:n_levels]]
