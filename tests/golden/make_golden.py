"""Generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Sources of the expected values
  * chainer path  -- chainer_maskrcnn/functions/roi_align/roi_align_2d.py
                     ROIAlign2D.forward_cpu / backward_cpu, imported as they are
                     under a stub `chainer` (oracle/reference_loader.py);
  * caffe2 path   -- the reference's caffe2_roi_align.cpp compiled into
                     oracle/_ref (forward only; there is no reference backward
                     for sampling_ratio > 1, so none is stored);
  * levels        -- map_rois_to_fpn_levels executed from the reference file.
The GPU box has no /root/reference; tests there compare against these files.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from oracle import reference_loader as ref  # noqa: E402
import synth  # noqa: E402


def reference_fixture():
    """The reference test's own fixture (test_roi_align_2d.py:17-37) with a fixed
    seed and 16 channels instead of 256 to keep the file small."""
    rng = np.random.RandomState(1234)
    N, C, H, W = 3, 16, 12, 8
    x = np.arange(N * C * H * W, dtype=np.float32).reshape(N, C, H, W)
    rng.shuffle(x)
    x = (2 * x / x.size - 1).astype(np.float32)
    rois = np.array([[0, 1, 1, 6, 6], [2, 6, 2, 7, 11], [1, 3, 1, 5, 10], [0, 3, 3, 3, 3]],
                    dtype=np.float32)
    rois = np.tile(rois, (15, 1))
    outh, outw, scale = 5, 7, 0.6
    gy = rng.uniform(-1, 1, (rois.shape[0], C, outh, outw)).astype(np.float32)
    y = ref.reference_forward(x, rois, outh, outw, scale)
    gx = ref.reference_backward(gy, x, rois, outh, outw, scale)
    out = dict(x=x, rois=rois, outh=outh, outw=outw, scale=np.float32(scale), gy=gy, y=y, gx=gx)
    for S in (1, 2, 3):
        out["y_caffe2_s%d" % S] = oracle.ref_caffe2_forward(x, rois, outh, outw, scale, S)
    return out


def fpn_small():
    """2 images of 160x192 px, P2-P5, 4 channels, 28 RoIs/image, box 7x7 and mask
    14x14, pooled level by level with the reference op (what the heads' loops
    compute, fpn_roi_mask_head.py:57-63,74-78) and the reference level mapper."""
    rng = np.random.RandomState(77)
    n_img, C, Himg, Wimg, L = 2, 4, 160, 192, 4
    feats = synth.make_pyramid(rng, n_img, C, Himg, Wimg, L)
    rois = synth.make_rois(rng, n_img, 28, Himg, Wimg, size_range=(6.0, 220.0))
    mapper = ref.reference_level_mapper()
    levels_f = mapper(rois[:, 1:])
    levels = np.clip(levels_f, 0, L - 1).astype(np.int32)   # maskrcnn.py:141, head :58
    scales = [1.0 / s for s in synth.STRIDES[:L]]
    rois_xy = rois[:, [0, 2, 1, 4, 3]]                        # roi_align_2d_yx.py:5
    out = dict(rois=rois, levels_f32=levels_f.astype(np.float32), levels=levels,
               scales=np.array(scales, np.float32))
    for l, f in enumerate(feats):
        out["feat%d" % l] = f
    for P in (7, 14):
        gy = synth.make_gy(rng, rois.shape[0], C, P)
        y = np.zeros((rois.shape[0], C, P, P), np.float32)
        out["gy%d" % P] = gy
        for l in range(L):
            sel = np.nonzero(levels == l)[0]
            gx = np.zeros_like(feats[l])
            if sel.size:
                y[sel] = ref.reference_forward(feats[l], rois_xy[sel], P, P, scales[l])
                gx = ref.reference_backward(np.ascontiguousarray(gy[sel]), feats[l],
                                            rois_xy[sel], P, P, scales[l])
            out["gx%d_l%d" % (P, l)] = gx
        out["y%d" % P] = y
        y2 = np.zeros_like(y)
        for l in range(L):
            sel = np.nonzero(levels == l)[0]
            if sel.size:
                y2[sel] = oracle.ref_caffe2_forward(feats[l], rois_xy[sel], P, P, scales[l], 2)
        out["y%d_caffe2_s2" % P] = y2
    return out


def level_vectors():
    """Boxes straddling every level boundary plus random ones, with the reference
    mapper's float32 output."""
    rng = np.random.RandomState(5)
    mapper = ref.reference_level_mapper()
    boxes = []
    for side in (28.0, 56.0, 112.0, 224.0, 448.0):
        for d in np.linspace(-0.01, 0.01, 41):
            h = side + d
            boxes.append([10.0, 20.0, 10.0 + h, 20.0 + h])
            boxes.append([0.0, 0.0, h * 2.0, h / 2.0])
    boxes = np.array(boxes, np.float32)
    rnd = np.zeros((4000, 4), np.float32)
    rnd[:, :2] = rng.uniform(0, 800, (4000, 2))
    rnd[:, 2:] = rnd[:, :2] + np.exp(rng.uniform(np.log(2), np.log(900), (4000, 2)))
    boxes = np.concatenate([boxes, rnd.astype(np.float32),
                            np.array([[5, 5, 5, 5], [5, 5, 5, 50]], np.float32)])
    return dict(boxes=boxes, levels=mapper(boxes).astype(np.float32),
                levels_kmax3=mapper(boxes, 0, 3).astype(np.float32),
                thresholds=oracle.level_area_thresholds())


def sweep(n_cases=8):
    """Random small single-level cases (odd channel counts, rectangular outputs, maps
    narrower than the vectorised path needs, degenerate boxes): the reference NumPy op
    forward/backward on in-bounds boxes, and the reference C++ forward at a random
    sampling grid on the same boxes pushed over the borders."""
    out = dict(n_cases=np.int32(n_cases))
    for i in range(n_cases):
        rng = np.random.RandomState(4200 + i)
        N, C = int(rng.randint(1, 3)), int(rng.randint(1, 7))
        H, W = int(rng.randint(5, 33)), int(rng.randint(5, 33))
        scale = float(rng.choice([1.0, 0.5, 0.25, 0.6]))
        outh, outw = int(rng.randint(1, 15)), int(rng.randint(1, 15))
        R = int(rng.randint(2, 13))
        x = rng.standard_normal((N, C, H, W)).astype(np.float32)
        b = rng.randint(0, N, R).astype(np.float32)
        x1 = rng.uniform(0, (W - 1) / scale, R)
        y1 = rng.uniform(0, (H - 1) / scale, R)
        x2 = x1 + rng.uniform(0, 1, R) * ((W - 1) / scale - x1)
        y2 = y1 + rng.uniform(0, 1, R) * ((H - 1) / scale - y1)
        x2[0], y2[0] = x1[0], y1[0]                                   # a degenerate box
        rois = np.stack([b, x1, y1, x2, y2], 1).astype(np.float32)
        gy = rng.uniform(-1, 1, (R, C, outh, outw)).astype(np.float32)
        S = int(rng.choice([0, 1, 2, 3]))
        rois_c2 = rois.copy()
        k = rng.rand(R) < 0.4
        rois_c2[k, 1:3] -= rng.uniform(0, 10, (int(k.sum()), 2)).astype(np.float32)
        rois_c2[k, 3:5] += rng.uniform(0, 10, (int(k.sum()), 2)).astype(np.float32)
        pre = "c%d_" % i
        out.update({pre + "x": x, pre + "rois": rois, pre + "gy": gy,
                    pre + "geom": np.array([outh, outw, S], np.int32), pre + "scale": np.float32(scale),
                    pre + "y": ref.reference_forward(x, rois, outh, outw, scale),
                    pre + "gx": ref.reference_backward(gy, x, rois, outh, outw, scale),
                    pre + "rois_c2": rois_c2,
                    pre + "y_c2": oracle.ref_caffe2_forward(x, rois_c2, outh, outw, scale, S)})
    return out


def main():
    assert ref.available(), "needs the reference tree"
    assert oracle.have_ref() or (oracle.build() or oracle.have_ref())
    np.savez_compressed(os.path.join(HERE, "reference_fixture.npz"), **reference_fixture())
    np.savez_compressed(os.path.join(HERE, "fpn_small.npz"), **fpn_small())
    np.savez_compressed(os.path.join(HERE, "levels.npz"), **level_vectors())
    np.savez_compressed(os.path.join(HERE, "sweep.npz"), **sweep())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
