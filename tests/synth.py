"""Seeded synthetic inputs for tests and bench.py.

Features ~ N(0,1) float32 in the reference's NCHW layout; RoIs with sqrt(area)
log-uniform in [16, 512] px, aspect h/w log-uniform, clipped to the image and
placed uniformly inside it, stored as (R,5) float32 [image, y1, x1, y2, x2]
(the heads' indices_and_rois, fpn_maskrcnn_train_chain.py:73-78); gy ~ U(-1,1).
"""
import math

import numpy as np

STRIDES = [4, 8, 16, 32, 64]


def pyramid_shapes(n_images, channels, height, width, n_levels):
    return [(n_images, channels, math.ceil(height / s), math.ceil(width / s))
            for s in STRIDES[:n_levels]]


def make_rois(rng, n_images, rois_per_image, height, width, size_range=(16.0, 512.0),
              aspect_range=(0.5, 2.0)):
    out = []
    for b in range(n_images):
        s = np.exp(rng.uniform(math.log(size_range[0]), math.log(size_range[1]), rois_per_image))
        a = np.exp(rng.uniform(math.log(aspect_range[0]), math.log(aspect_range[1]), rois_per_image))
        h = np.minimum(s * np.sqrt(a), height)
        w = np.minimum(s / np.sqrt(a), width)
        y1 = rng.uniform(0.0, 1.0, rois_per_image) * (height - h)
        x1 = rng.uniform(0.0, 1.0, rois_per_image) * (width - w)
        r = np.stack([np.full(rois_per_image, b, np.float64), y1, x1, y1 + h, x1 + w], axis=1)
        out.append(r)
    rois = np.concatenate(out).astype(np.float32)
    # float32 rounding must not push a box past the image
    rois[:, 3] = np.minimum(rois[:, 3], np.float32(height))
    rois[:, 4] = np.minimum(rois[:, 4], np.float32(width))
    return rois


def make_pyramid(rng, n_images, channels, height, width, n_levels):
    return [rng.standard_normal(s).astype(np.float32)
            for s in pyramid_shapes(n_images, channels, height, width, n_levels)]


def make_gy(rng, n_rois, channels, out_size):
    return rng.uniform(-1.0, 1.0, (n_rois, channels, out_size, out_size)).astype(np.float32)


# BASELINE.json configs (index = position in "configs")
CONFIGS = {
    0: dict(name="cfg0_box7_1img_800x800_512rois", n_images=1, height=800, width=800,
            rois_per_image=512, out_sizes=[7], n_levels=4, channels=256, aspect=(0.5, 2.0)),
    1: dict(name="cfg1_mask14_2img_800x1333_2048rois", n_images=2, height=800, width=1333,
            rois_per_image=2048, out_sizes=[14], n_levels=4, channels=256, aspect=(0.5, 2.0)),
    2: dict(name="cfg2_keypoint14_4img_800x1333_512rois", n_images=4, height=800, width=1333,
            rois_per_image=512, out_sizes=[14], n_levels=4, channels=256, aspect=(1.5, 3.0)),
    3: dict(name="cfg3_box7_mask14_16img_800x1333_1000rois", n_images=16, height=800, width=1333,
            rois_per_image=1000, out_sizes=[7, 14], n_levels=4, channels=256, aspect=(0.5, 2.0)),
    # what ONE of 8 GPUs holds when configs[3] is sharded by image (tuning aid: same shapes, 2 images)
    13: dict(name="cfg3_one_eighth_2img_1000rois", n_images=2, height=800, width=1333,
             rois_per_image=1000, out_sizes=[7, 14], n_levels=4, channels=256, aspect=(0.5, 2.0)),
    # configs[1] with half the channels (tuning aid: a bin row of gy is one slab wide)
    21: dict(name="cfg1_mask14_128ch", n_images=2, height=800, width=1333,
             rois_per_image=2048, out_sizes=[14], n_levels=4, channels=128, aspect=(0.5, 2.0)),
}


def window_cells_touched(rois, levels, shapes, scales, samples_per_side):
    """U of SURVEY.md 8(d): number of distinct (image, level, y, x) cells whose
    value at least one RoI reads.  A RoI's taps are the outer product of the rows
    and the columns its samples touch: from floor(first sample) to
    floor(last sample) + 1, samples_per_side = pooled size * sampling_ratio of the
    finest head (its sample grid has the outermost samples)."""
    total = 0
    for l, shp in enumerate(shapes):
        n, _, H, W = shp
        mask = np.zeros((n, H, W), dtype=bool)
        sel = np.nonzero(levels == l)[0]
        sc = scales[l]
        for r in sel:
            b, y1, x1, y2, x2 = (float(v) for v in rois[r])
            ys, xs = y1 * sc, x1 * sc
            rh, rw = max(y2 * sc - ys, 1.0), max(x2 * sc - xs, 1.0)
            hy, hx = 0.5 * rh / samples_per_side, 0.5 * rw / samples_per_side
            ya = min(max(int(math.floor(ys + hy)), 0), H - 1)
            yb = min(max(int(math.floor(ys + rh - hy)), 0) + 1, H - 1)
            xa = min(max(int(math.floor(xs + hx)), 0), W - 1)
            xb = min(max(int(math.floor(xs + rw - hx)), 0) + 1, W - 1)
            mask[int(b), ya:yb + 1, xa:xb + 1] = True
        total += int(mask.sum())
    return total
