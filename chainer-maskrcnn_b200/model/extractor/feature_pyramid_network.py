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


# ---------------------------------------------------------------------------
# cfg 5: the extractor itself, restated on torch so that a whole
# training step can run around the pooling path.  Dense cuDNN work -- none of it is
# part of librpool_b200.so.
# ---------------------------------------------------------------------------
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402


class _Bottleneck(nn.Module):
    """ResNet-50 building block as chainer's ResNet50Layers lays it out (Caffe
    style: the stride sits on the first 1x1 convolution)."""

    def __init__(self, cin, mid, cout, stride, project):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, mid, 1, stride, 0, bias=False)
        self.bn1 = nn.BatchNorm2d(mid)
        self.conv2 = nn.Conv2d(mid, mid, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(mid)
        self.conv3 = nn.Conv2d(mid, cout, 1, 1, 0, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        self.proj = None
        if project:
            self.proj = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, 0, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        h = F.relu(self.bn1(self.conv1(x)))
        h = F.relu(self.bn2(self.conv2(h)))
        h = self.bn3(self.conv3(h))
        return F.relu(h + (x if self.proj is None else self.proj(x)))


def _stage(cin, mid, cout, n, stride):
    blocks = [_Bottleneck(cin, mid, cout, stride, True)]
    blocks += [_Bottleneck(cout, mid, cout, 1, False) for _ in range(n - 1)]
    return nn.Sequential(*blocks)


def _unpool2(x, size):
    """F.unpooling_2d(x, ksize=2, outsize=size) (feature_pyramid_network.py:58-66):
    every cell repeated 2 x 2, cropped to the lateral map's size."""
    x = x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    return x[:, :, :size[0], :size[1]]


class FeaturePyramidNetwork(nn.Module):
    """ResNet-50 bottom-up + top-down pyramid, p2..p6 fine to coarse
    (chainer_maskrcnn/model/extractor/feature_pyramid_network.py:18-71).
    ``width`` scales the backbone (64 = ResNet-50; tests use a thin one)."""
    feat_strides = feat_strides
    spatial_scales = spatial_scales
    anchor_base = 16
    anchor_sizes = [32, 64, 128, 256, 512]
    anchor_scales = [s / 16. for s in anchor_sizes]

    def __init__(self, width=64, blocks=(3, 4, 6, 3), out_channels=256):
        super().__init__()
        w = width
        self.conv1 = nn.Conv2d(3, w, 7, 2, 3)
        self.bn1 = nn.BatchNorm2d(w)
        self.res2 = _stage(w, w, 4 * w, blocks[0], 1)
        self.res3 = _stage(4 * w, 2 * w, 8 * w, blocks[1], 2)
        self.res4 = _stage(8 * w, 4 * w, 16 * w, blocks[2], 2)
        self.res5 = _stage(16 * w, 8 * w, 32 * w, blocks[3], 2)
        c = out_channels
        self.toplayer = nn.Conv2d(32 * w, c, 1)
        self.conv_p4 = nn.Conv2d(c, c, 3, 1, 1)
        self.conv_p3 = nn.Conv2d(c, c, 3, 1, 1)
        self.conv_p2 = nn.Conv2d(c, c, 3, 1, 1)
        self.conv_p6 = nn.Conv2d(c, c, 1, 2, 0)
        self.lat_p4 = nn.Conv2d(16 * w, c, 1)
        self.lat_p3 = nn.Conv2d(8 * w, c, 1)
        self.lat_p2 = nn.Conv2d(4 * w, c, 1)

    def forward(self, x):
        h = F.relu(self.bn1(self.conv1(x)))
        h = F.max_pool2d(h, 2, ceil_mode=True)      # max_pooling_2d(ksize=2): cover_all
        c2 = self.res2(h)
        c3 = self.res3(c2)
        c4 = self.res4(c3)
        c5 = self.res5(c4)
        p5 = self.toplayer(c5)
        p4 = self.conv_p4(_unpool2(p5, c4.shape[2:]) + self.lat_p4(c4))
        p3 = self.conv_p3(_unpool2(p4, c3.shape[2:]) + self.lat_p3(c3))
        p2 = self.conv_p2(_unpool2(p3, c2.shape[2:]) + self.lat_p2(c2))
        p6 = self.conv_p6(p5)
        return p2, p3, p4, p5, p6
