"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on
the same seeded inputs, against the committed golden fixtures, and at
BASELINE.json's full sizes.

Tolerances (BASELINE.json north_star), metric max|a-b| / max|b|:
    forward  <= 1e-5      backward <= 1e-4 (atomic reordering)
    level assignment and output<->RoI order: bit-exact
The generic kernel path replays the reference's operation order and is
checked bit-for-bit in the forward direction.
"""
import os

import numpy as np
import pytest
import torch

import oracle
import synth
from chainer_maskrcnn_b200 import _engine, _lib

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-5
BWD_TOL = 1e-4
PATHS = {"auto": _lib.PATH_AUTO, "generic": _lib.PATH_GENERIC, "table": _lib.PATH_TABLE}


def dev(a, channels_last=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    if channels_last and t.dim() == 4:
        t = t.contiguous(memory_format=torch.channels_last)
    return t


def host(t):
    return np.ascontiguousarray(t.detach().cpu().numpy())


def run_fused(feats, rois_yx, levels, scales, out_sizes, S=1, mode=None, channels_last=True,
              gys=None, levels_dtype=np.int32, deterministic=False, options=None):
    f = [dev(x, channels_last) for x in feats]
    lv = None if levels is None else dev(np.asarray(levels).astype(levels_dtype))
    outs, plan = _engine.forward(f, dev(rois_yx), lv, scales, out_sizes, sampling_ratio=S,
                                 coord_mode=mode, roi_format=_lib.ROI_YX, options=options)
    for o in outs:
        assert o.is_contiguous(memory_format=torch.channels_last) or o.numel() == 0 or o.shape[1] == 1 \
            or o.shape[1] % 4 != 0        # (a channel count the shim padded: a view of the padded result)
    grads = None
    if gys is not None:
        grads = [host(g) for g in _engine.backward(plan, [dev(g) for g in gys],
                                                   deterministic=deterministic)]
    torch.cuda.synchronize()
    return [host(o) for o in outs], grads, plan


def oracle_fused(feats, rois_yx, levels, scales, out_sizes, S, mode_name, gys=None):
    outs = [oracle.fpn_forward(feats, rois_yx, levels, scales, P, mode_name, S, threads=8)
            for P in out_sizes]
    grads = None
    if gys is not None:
        grads = [np.zeros_like(f) for f in feats]
        for g in gys:
            part = oracle.fpn_backward(g, [f.shape for f in feats], rois_yx, levels, scales,
                                       mode_name, S, threads=8)
            for l in range(len(feats)):
                grads[l] += part[l]
    return outs, grads


def make_case(seed, n_img=2, C=64, Himg=256, Wimg=320, L=4, per_img=150, size=(8.0, 300.0),
              aspect=(0.5, 2.0)):
    rng = np.random.RandomState(seed)
    feats = synth.make_pyramid(rng, n_img, C, Himg, Wimg, L)
    rois = synth.make_rois(rng, n_img, per_img, Himg, Wimg, size_range=size, aspect_range=aspect)
    rng.shuffle(rois)  # interleave images: the schedule must not depend on input order
    levels = oracle.levels_for_pyramid(rois[:, 1:], L)
    scales = [1.0 / s for s in synth.STRIDES[:L]]
    return rng, feats, rois, levels, scales


# ---------------------------------------------------------------------------
# golden fixtures generated from the unmodified reference
# ---------------------------------------------------------------------------
def test_golden_reference_fixture_single_level_op(golden_dir):
    from chainer_maskrcnn_b200 import ROIAlign2D
    d = np.load(os.path.join(golden_dir, "reference_fixture.npz"))
    outh, outw, scale = int(d["outh"]), int(d["outw"]), float(d["scale"])
    x, rois, gy = dev(d["x"]), dev(d["rois"]), dev(d["gy"])
    for path in ("auto", "generic", "table"):
        opt = dict(force_path=PATHS[path])
        f = ROIAlign2D(outh, outw, scale, options=opt)
        (y,) = f.forward_gpu((x, rois))
        assert y.dtype == torch.float32 and tuple(y.shape) == tuple(d["gy"].shape)
        assert oracle.rel_err(host(y), d["y"]) <= FWD_TOL, path
        if path == "generic":
            # C = 16 is a multiple of 4 but the generic path was forced: reference op order
            assert np.array_equal(host(y), d["y"])
        gx, none = f.backward_gpu((None, rois), (gy,))
        assert none is None
        assert oracle.rel_err(host(gx), d["gx"]) <= BWD_TOL, path
        for S in (1, 2, 3):
            (y2,), _ = _engine.forward([x], rois, None, [scale], [(outh, outw)], S,
                                       _lib.COORD_CAFFE2, _lib.ROI_XY, options=opt)
            assert oracle.rel_err(host(y2), d["y_caffe2_s%d" % S]) <= FWD_TOL, (path, S)
            if path == "generic":
                assert np.array_equal(host(y2), d["y_caffe2_s%d" % S])


def test_golden_fpn_small_fused_two_heads(golden_dir):
    d = np.load(os.path.join(golden_dir, "fpn_small.npz"))
    feats = [d["feat%d" % l] for l in range(4)]
    scales = [float(s) for s in d["scales"]]
    for levels in (None, d["levels"], d["levels_f32"]):
        ldt = np.float32 if levels is not None and levels.dtype == np.float32 else np.int32
        for cl in (True, False):
            outs, grads, plan = run_fused(feats, d["rois"], levels, scales, [7, 14], 1, None, cl,
                                          gys=[d["gy7"], d["gy14"]], levels_dtype=ldt)
            lv, order = _engine.read_plan(plan)
            assert np.array_equal(lv, d["levels"])             # bit-exact level assignment
            assert oracle.rel_err(outs[0], d["y7"]) <= FWD_TOL
            assert oracle.rel_err(outs[1], d["y14"]) <= FWD_TOL
            for l in range(4):
                want = d["gx7_l%d" % l] + d["gx14_l%d" % l]
                assert oracle.rel_err(grads[l], want) <= BWD_TOL, l
    outs, _, _ = run_fused(feats, d["rois"], d["levels"], scales, [7, 14], 2)
    assert oracle.rel_err(outs[0], d["y7_caffe2_s2"]) <= FWD_TOL
    assert oracle.rel_err(outs[1], d["y14_caffe2_s2"]) <= FWD_TOL


def test_golden_sweep_single_level_op(golden_dir):
    # random small cases from the reference itself (odd channel counts, rectangular outputs,
    # narrow maps, degenerate boxes; C++ forward with boxes over the borders)
    from chainer_maskrcnn_b200 import ROIAlign2D
    d = np.load(os.path.join(golden_dir, "sweep.npz"))
    for i in range(int(d["n_cases"])):
        g = lambda k: d["c%d_%s" % (i, k)]
        outh, outw, S = (int(v) for v in g("geom"))
        scale = float(g("scale"))
        x, rois = dev(g("x")), dev(g("rois"))
        f = ROIAlign2D(outh, outw, scale)
        (y,) = f.forward_gpu((x, rois))
        assert tuple(y.shape) == g("y").shape
        assert oracle.rel_err(host(y), g("y")) <= FWD_TOL, i
        gx, _ = f.backward_gpu((None, rois), (dev(g("gy")),))
        assert tuple(gx.shape) == g("x").shape
        assert oracle.rel_err(host(gx), g("gx")) <= BWD_TOL, i
        (y2,), _ = _engine.forward([x], dev(g("rois_c2")), None, [scale], [(outh, outw)], S,
                                   _lib.COORD_CAFFE2, _lib.ROI_XY)
        assert oracle.rel_err(host(y2), g("y_c2")) <= FWD_TOL, (i, S)


def test_golden_levels_bit_exact(golden_dir):
    d = np.load(os.path.join(golden_dir, "levels.npz"))
    got = host(_engine.assign_levels(dev(d["boxes"])))
    assert got.dtype == np.float32 and np.array_equal(got, d["levels"])
    got3 = host(_engine.assign_levels(dev(d["boxes"]), k_max=3))
    assert np.array_equal(got3, d["levels_kmax3"])
    from chainer_maskrcnn_b200 import map_rois_to_fpn_levels
    assert np.array_equal(host(map_rois_to_fpn_levels(dev(d["boxes"]))), d["levels"])
    assert np.array_equal(map_rois_to_fpn_levels(d["boxes"]), d["levels"])   # host arrays in/out


def test_levels_random_million_bit_exact():
    rng = np.random.RandomState(11)
    n = 1_000_000
    b = np.zeros((n, 4), np.float32)
    b[:, :2] = rng.uniform(0, 1400, (n, 2))
    b[:, 2:] = b[:, :2] + np.exp(rng.uniform(np.log(0.5), np.log(1400), (n, 2))).astype(np.float32)
    want = oracle.map_rois_to_fpn_levels(b)
    got = host(_engine.assign_levels(dev(b)))
    assert np.array_equal(got, want)
    gi = host(_engine.assign_levels(dev(b), k_cap=3, as_int=True))
    assert gi.dtype == np.int32 and np.array_equal(gi, np.clip(want, 0, 3).astype(np.int32))


# ---------------------------------------------------------------------------
# seeded random cases against the oracle, every kernel path
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("path,prefetch", [("auto", {}), ("generic", dict(prefetch_rows=-1)),
                                           ("table", dict(prefetch_rois=1)), ("table", dict(prefetch_rois=8)),
                                           ("table", dict(prefetch_rows=1)), ("table", dict(prefetch_rows=-1))])
@pytest.mark.parametrize("mode_name,S", [("chainer", 1), ("caffe2", 1), ("caffe2", 2), ("caffe2", 3)])
def test_fused_vs_oracle(path, prefetch, mode_name, S):
    rng, feats, rois, levels, scales = make_case(seed=S * 7 + len(path))
    opt = dict(force_path=PATHS[path], **prefetch)
    mode = _lib.COORD_CHAINER if mode_name == "chainer" else _lib.COORD_CAFFE2
    sizes = [7, 14]
    gys = [synth.make_gy(rng, rois.shape[0], feats[0].shape[1], P) for P in sizes]
    outs, grads, plan = run_fused(feats, rois, levels, scales, sizes, S, mode, gys=gys, options=opt)
    want, wgrads = oracle_fused(feats, rois, levels, scales, sizes, S, mode_name, gys)
    for o, w in zip(outs, want):
        assert oracle.rel_err(o, w) <= FWD_TOL
        if path == "generic":
            assert np.array_equal(o, w)          # reference operation order
    for g, w in zip(grads, wgrads):
        assert oracle.rel_err(g, w) <= BWD_TOL


@pytest.mark.parametrize("threads", [32, 64, 224, 256])
def test_block_sizes(threads):
    rng, feats, rois, levels, scales = make_case(seed=3, C=128, per_img=60)
    gys = [synth.make_gy(rng, rois.shape[0], 128, 14)]
    outs, grads, _ = run_fused(feats, rois, levels, scales, [14], 2, gys=gys,
                               options=dict(cta_threads=threads))
    want, wgrads = oracle_fused(feats, rois, levels, scales, [14], 2, "caffe2", gys)
    assert oracle.rel_err(outs[0], want[0]) <= FWD_TOL
    for g, w in zip(grads, wgrads):
        assert oracle.rel_err(g, w) <= BWD_TOL


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("mode_name,S,sizes,C,size", [
    ("caffe2", 2, [14], 256, (8.0, 300.0)), ("chainer", 1, [14], 256, (8.0, 300.0)),
    ("caffe2", 2, [7, 14], 128, (8.0, 300.0)), ("caffe2", 0, [7], 384, (8.0, 300.0)),
    ("caffe2", 2, [14], 256, (1.0, 12.0)),      # tiny RoIs: many bin rows on one window row
    ("caffe2", 1, [16], 128, (8.0, 300.0)), ("caffe2", 2, [14], 64, (8.0, 300.0))])   # (C = 64: rows kernel either way)
def test_backward_variants(variant, mode_name, S, sizes, C, size):
    """opt.backward_variant: the rows kernel (a CTA per RoI) and the staged kernel (persistent CTAs,
    gy through shared memory by bulk copies) both match the oracle."""
    rng, feats, rois, levels, scales = make_case(seed=51 + S, C=C, per_img=400, size=size)
    mode = _lib.COORD_CHAINER if mode_name == "chainer" else _lib.COORD_CAFFE2
    gys = [synth.make_gy(rng, rois.shape[0], C, P) for P in sizes]
    _, grads, _ = run_fused(feats, rois, levels, scales, sizes, S, mode, gys=gys,
                            options=dict(backward_variant=variant))
    _, want = oracle_fused(feats, rois, levels, scales, sizes, S, mode_name, gys)
    for g, w in zip(grads, want):
        assert oracle.rel_err(g, w) <= BWD_TOL


def test_staged_backward_few_rois_and_bad_rois():
    """Fewer RoIs than SMs, a RoI of an image outside the batch, one RoI only."""
    rng, feats, rois, levels, scales = make_case(seed=61, C=128, per_img=20)
    for n in (1, 7, 40):
        r2 = rois[:n].copy()
        if n > 3:
            r2[3, 0] = 9
        lv = levels[:n]
        gy = synth.make_gy(rng, n, 128, 14)
        _, a, _ = run_fused(feats, r2, lv, scales, [14], 2, gys=[gy], options=dict(backward_variant=2))
        _, b, _ = run_fused(feats, r2, lv, scales, [14], 2, gys=[gy], options=dict(backward_variant=1))
        for x, y in zip(a, b):
            assert oracle.rel_err(x, y) <= BWD_TOL


@pytest.mark.parametrize("split", [0, 1])
@pytest.mark.parametrize("mode_name,S", [("chainer", 1), ("caffe2", 2)])
def test_two_pooled_sizes_backward_fused_or_one_launch_per_size(split, mode_name, S):
    # one backward launch per pooled size (default) or, opt.fuse_heads_backward, one for both
    rng, feats, rois, levels, scales = make_case(31, per_img=120)
    C = feats[0].shape[1]
    gys = [synth.make_gy(rng, rois.shape[0], C, 7), synth.make_gy(rng, rois.shape[0], C, 14)]
    outs, grads, plan = run_fused(feats, rois, None, scales, [7, 14], S=S, gys=gys,
                                  options=dict(fuse_heads_backward=1 - split))
    g_cl = [dev(g, True) for g in gys]
    n0 = _lib.launch_count()
    _engine.backward(plan, g_cl)
    launches = _lib.launch_count() - n0
    want, want_g = oracle_fused(feats, rois, levels, scales, [7, 14], S, mode_name, gys)
    for o, w in zip(outs, want):
        assert oracle.rel_err(o, w) <= FWD_TOL
    for g, w in zip(grads, want_g):
        assert oracle.rel_err(g, w) <= BWD_TOL
    assert launches == 1 + (2 if split else 1)   # zero fill + backward launch(es)


def test_rectangular_output_and_odd_channels():
    # C = 6 is not a multiple of 4 -> generic kernel path when handed to the library as it is
    # (bit-equal forward); the shim's default pads it to 8 zero-extended channels and takes the
    # vectorised path (within tolerance).  (5,7) output as in the reference test
    rng, feats, rois, levels, scales = make_case(seed=5, C=6, per_img=40)
    gys = [rng.uniform(-1, 1, (rois.shape[0], 6, 5, 7)).astype(np.float32)]
    f = [dev(x) for x in feats]
    outs_p, plan_p = _engine.forward(f, dev(rois), dev(levels), scales, [(5, 7)])
    grads_p = _engine.backward(plan_p, [dev(gys[0])])
    assert plan_p.cpad == 8 and tuple(outs_p[0].shape) == (rois.shape[0], 6, 5, 7)
    assert all(tuple(g.shape) == x.shape for g, x in zip(grads_p, feats))
    outs, plan = _engine.forward(f, dev(rois), dev(levels), scales, [(5, 7)], pad_channels=False)
    grads = _engine.backward(plan, [dev(gys[0])])
    assert plan.cpad == 6
    rois_xy = oracle.roi_yx_to_xy(rois)
    want = np.zeros((rois.shape[0], 6, 5, 7), np.float32)
    for l in range(4):
        sel = np.nonzero(levels == l)[0]
        want[sel] = oracle.forward_chainer(feats[l], rois_xy[sel], 5, 7, scales[l])
        wg = oracle.backward_chainer(gys[0][sel], rois_xy[sel], feats[l].shape, scales[l])
        assert oracle.rel_err(host(grads[l]), wg) <= BWD_TOL
        assert oracle.rel_err(host(grads_p[l]), wg) <= BWD_TOL
    assert np.array_equal(host(outs[0]), want)
    assert oracle.rel_err(host(outs_p[0]), want) <= FWD_TOL
    # gradients into caller-provided buffers, overwrite and accumulate, through the padded path
    mine = [torch.zeros_like(dev(x, True)) for x in feats]
    _engine.backward(plan_p, [dev(gys[0])], out=mine)
    _engine.backward(plan_p, [dev(gys[0])], out=mine, accumulate=True)
    for l in range(4):
        assert oracle.rel_err(host(mine[l]), 2 * host(grads[l])) <= BWD_TOL
    # same with C = 8 (fast path), rectangular
    rng, feats, rois, levels, scales = make_case(seed=6, C=8, per_img=40)
    outs, plan = _engine.forward([dev(x) for x in feats], dev(rois), dev(levels), scales, [(5, 7)])
    rois_xy = oracle.roi_yx_to_xy(rois)
    for l in range(4):
        sel = np.nonzero(levels == l)[0]
        w = oracle.forward_chainer(feats[l], rois_xy[sel], 5, 7, scales[l])
        assert oracle.rel_err(host(outs[0])[sel], w) <= FWD_TOL


def test_large_pooled_size_and_adaptive_sampling():
    rng, feats, rois, levels, scales = make_case(seed=8, C=8, per_img=30)
    # P = 40 > fast-path table size -> generic path, which keeps the reference's
    # operation order and is therefore bit-equal.  S = 0 (caffe2 adaptive grid,
    # ceil(bin size) samples per side) takes the table path whenever a bin's
    # footprint fits it: within tolerance there, bit-equal when forced generic.
    for P, S in ((40, 2), (7, 0), (3, 0)):
        want, _ = oracle_fused(feats, rois, levels, scales, [P], S, "caffe2")
        outs, _, _ = run_fused(feats, rois, levels, scales, [P], S, _lib.COORD_CAFFE2)
        if P > 32:
            assert np.array_equal(outs[0], want[0]), (P, S)
        else:
            assert oracle.rel_err(outs[0], want[0]) <= FWD_TOL, (P, S)
        outs, _, _ = run_fused(feats, rois, levels, scales, [P], S, _lib.COORD_CAFFE2,
                               options=dict(force_path=_lib.PATH_GENERIC))
        assert np.array_equal(outs[0], want[0]), (P, S, "generic")
    gy = synth.make_gy(rng, rois.shape[0], 8, 7)
    _, grads, _ = run_fused(feats, rois, levels, scales, [7], 0, _lib.COORD_CAFFE2, gys=[gy])
    _, wg = oracle_fused(feats, rois, levels, scales, [7], 0, "caffe2", [gy])
    for g, w in zip(grads, wg):
        assert oracle.rel_err(g, w) <= BWD_TOL


def test_wide_bins_single_level():
    """The reference micro-benchmark shape (test_performance.py:19-25): 200 px RoIs
    on a stride-1 map, 14x14 -> bins 14 cells wide (sliding window must jump)."""
    rng = np.random.RandomState(9)
    x = rng.standard_normal((1, 16, 224, 224)).astype(np.float32)
    rois = np.array([[0, 0, 0, 200, 200], [0, 10.5, 3.25, 190, 222], [0, 100, 100, 101, 180]], np.float32)
    from chainer_maskrcnn_b200 import roi_align_2d
    for S in (1, 2):
        y = roi_align_2d(dev(x), dev(rois), 14, 14, 1.0, sampling_ratio=S)
        w = oracle.forward_chainer(x, rois, 14, 14, 1.0) if S == 1 else \
            oracle.forward_caffe2(x, rois, 14, 14, 1.0, S)
        assert oracle.rel_err(host(y), w) <= FWD_TOL


def test_edge_cases_caffe2_borders_and_degenerate():
    rng = np.random.RandomState(10)
    x = rng.standard_normal((2, 8, 20, 27)).astype(np.float32)
    rois = np.array([[0, -8, -8, 10, 10], [1, 40, 30, 70, 50], [0, -30, 5, -20, 9],
                     [1, 10, 10, 10, 10], [0, 0, 0, 54, 40], [1, 53.9, 39.9, 54, 40],
                     [0, 25, 18, 80, 70]], np.float32)       # xy format
    gy = rng.uniform(-1, 1, (rois.shape[0], 8, 7, 7)).astype(np.float32)
    for path in ("auto", "generic", "table"):
        opt = dict(force_path=PATHS[path])
        for S in (1, 2):
            outs, plan = _engine.forward([dev(x)], dev(rois), None, [0.5], [7], S,
                                         _lib.COORD_CAFFE2, _lib.ROI_XY, options=opt)
            w = oracle.forward_caffe2(x, rois, 7, 7, 0.5, S)
            assert oracle.rel_err(host(outs[0]), w) <= FWD_TOL, (path, S)
            g = _engine.backward(plan, [dev(gy)])
            wg = oracle.backward_caffe2(gy, rois, x.shape, 0.5, S)
            assert oracle.rel_err(host(g[0]), wg) <= BWD_TOL, (path, S)
    # chainer mode: degenerate boxes (double-precision stride branch) and negative origins
    rois_c = np.array([[0, 3, 3, 3, 3], [1, 2.5, 4, 2.6, 9], [0, 1, 1, 1.2, 30], [1, -1.5, -0.75, 20, 12]],
                      np.float32)
    gyc = rng.uniform(-1, 1, (rois_c.shape[0], 8, 5, 7)).astype(np.float32)
    for path in ("auto", "generic"):
        outs, plan = _engine.forward([dev(x)], dev(rois_c), None, [0.5], [(5, 7)], 1,
                                     _lib.COORD_CHAINER, _lib.ROI_XY, options=dict(force_path=PATHS[path]))
        w = oracle.forward_chainer(x, rois_c, 5, 7, 0.5)
        assert oracle.rel_err(host(outs[0]), w) <= FWD_TOL
        if path == "generic":
            assert np.array_equal(host(outs[0]), w)
        g = _engine.backward(plan, [dev(gyc)])
        assert oracle.rel_err(host(g[0]), oracle.backward_chainer(gyc, rois_c, x.shape, 0.5)) <= BWD_TOL


def test_empty_invalid_batch_and_level_clipping():
    rng, feats, rois, levels, scales = make_case(seed=12, C=8, per_img=10)
    f = [dev(x) for x in feats]
    outs, plan = _engine.forward(f, dev(rois[:0]), dev(levels[:0]), scales, [7, 14])
    assert tuple(outs[0].shape) == (0, 8, 7, 7) and tuple(outs[1].shape) == (0, 8, 14, 14)
    grads = _engine.backward(plan, [dev(np.zeros((0, 8, 7, 7), np.float32)),
                                    dev(np.zeros((0, 8, 14, 14), np.float32))])
    assert all(float(g.abs().max()) == 0.0 for g in grads)      # zero-filled dense gradients
    # out-of-range levels are clipped to the pyramid (maskrcnn.py:141); a batch index
    # outside the tensor yields zeros and no gradient
    lv = levels.copy()
    lv[0], lv[1] = 9, -3
    r2 = rois.copy()
    r2[2, 0] = 7
    outs, plan = _engine.forward(f, dev(r2), dev(lv), scales, [7])
    want = oracle.fpn_forward(feats, np.delete(r2, 2, 0), np.clip(np.delete(lv, 2), 0, 3), scales, 7)
    got = host(outs[0])
    assert np.all(got[2] == 0)
    assert oracle.rel_err(np.delete(got, 2, 0), want) <= FWD_TOL
    # ... and both are flagged: the reference's NumPy path raises IndexError for such a RoI
    flags = _engine.status_flags(plan)
    assert flags & _lib.FLAG_BAD_BATCH and flags & _lib.FLAG_LEVEL_CLIPPED
    with pytest.raises(IndexError):
        _engine.check_rois(plan)
    _, clean = _engine.forward(f, dev(rois), dev(levels), scales, [7])
    assert _engine.status_flags(clean) == 0
    _engine.check_rois(clean)
    # levels in any integer / floating dtype (the heads cast with astype(int32))
    base = host(_engine.forward(f, dev(rois), dev(levels), scales, [7])[0][0])
    for t in (torch.int64, torch.int16, torch.uint8, torch.float64, torch.float16):
        o, _ = _engine.forward(f, dev(rois), torch.from_numpy(levels).cuda().to(t), scales, [7])
        assert np.array_equal(host(o[0]), base), t
    with pytest.raises(TypeError):
        _engine.forward(f, dev(rois), torch.from_numpy(levels).cuda() > 0, scales, [7])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two devices")
def test_tensors_on_a_device_that_is_not_current():
    """every launch (layout conversion included) must run on the tensors'
    device and its stream, whatever torch's current device is."""
    rng, feats, rois, levels, scales = make_case(seed=41, C=16, per_img=30)
    d1 = torch.device("cuda", 1)
    f1 = [torch.from_numpy(x).to(d1) for x in feats]                # NCHW: converted on device 1
    gy = synth.make_gy(rng, rois.shape[0], 16, 7)
    assert torch.cuda.current_device() == 0
    outs, plan = _engine.forward(f1, torch.from_numpy(rois).to(d1), None, scales, [7], sampling_ratio=2)
    grads = _engine.backward(plan, [torch.from_numpy(gy).to(d1)])
    assert torch.cuda.current_device() == 0 and outs[0].device == d1 and grads[0].device == d1
    torch.cuda.synchronize(d1)
    want, wg = oracle_fused(feats, rois, levels, scales, [7], 2, "caffe2", [gy])
    assert oracle.rel_err(host(outs[0]), want[0]) <= FWD_TOL
    for g, w in zip(grads, wg):
        assert oracle.rel_err(host(g), w) <= BWD_TOL
    with pytest.raises(ValueError):
        _engine.forward([x.cuda(0) for x in f1], torch.from_numpy(rois).to(d1), None, scales, [7])


@pytest.mark.parametrize("force_generic", [False, True])
def test_split_tail_of_a_launch_longer_than_two_waves(force_generic):
    """A forward launch of at least two waves serves the RoIs at its last launch places with two
    CTAs each (a share of the bin rows per CTA; one of them alone on the generic path): every row
    of both pooled sizes must still be written exactly once, for an odd RoI count too."""
    rng, feats, rois, levels, scales = make_case(seed=77, C=8, per_img=701)
    rois = rois[:-1]                       # 1401 RoIs: odd, more than two waves of 128-thread CTAs
    levels = levels[:-1]
    gys = [synth.make_gy(rng, rois.shape[0], 8, P) for P in (7, 14)]
    opts = dict(force_path=_lib.PATH_GENERIC) if force_generic else None
    outs, grads, plan = run_fused(feats, rois, None, scales, [7, 14], S=2, gys=gys, options=opts)
    lv, _ = _engine.read_plan(plan)
    assert np.array_equal(lv, levels)
    want, wg = oracle_fused(feats, rois, levels, scales, [7, 14], 2, "caffe2", gys)
    for o, w in zip(outs, want):
        assert oracle.rel_err(o, w) <= FWD_TOL
        if force_generic:
            assert np.array_equal(o, w)    # the generic path is bit-equal to the oracle
    for g, w in zip(grads, wg):
        assert oracle.rel_err(g, w) <= BWD_TOL


def test_schedule_is_stable_binning_and_rows_keep_input_order():
    rng, feats, rois, levels, scales = make_case(seed=13, C=8, per_img=500)
    for order_mode in (_lib.SCHED_INPUT, _lib.SCHED_DEFAULT, _lib.SCHED_LEVEL_DESC, _lib.SCHED_COARSE_FIRST):
        outs, _, plan = run_fused(feats, rois, None, scales, [7], options=dict(schedule=order_mode))
        lv, order = _engine.read_plan(plan)
        assert np.array_equal(lv, levels)
        img = rois[:, 0].astype(np.int64)
        if order_mode == _lib.SCHED_INPUT:
            want = np.arange(len(lv))
        elif order_mode == _lib.SCHED_DEFAULT:
            want = np.argsort(img * 4 + lv, kind="stable")
        elif order_mode == _lib.SCHED_LEVEL_DESC:
            want = np.argsort(img * 4 + (3 - lv), kind="stable")
        else:
            want = np.argsort((3 - lv) * (img.max() + 1) + img, kind="stable")
        assert np.array_equal(order, want.astype(np.int32))
        # row r of the output is RoI r whatever the schedule (fpn_roi_mask_head.py:59-63)
        ref_out = oracle.fpn_forward(feats, rois, levels, scales, 7)
        assert oracle.rel_err(outs[0], ref_out) <= FWD_TOL
        perm = rng.permutation(len(lv))
        outs_p, _, _ = run_fused(feats, rois[perm], None, scales, [7])
        assert np.array_equal(outs_p[0], outs[0][perm])


def test_forward_is_run_to_run_identical_and_backward_within_tolerance():
    rng, feats, rois, levels, scales = make_case(seed=14, per_img=200)
    gy = synth.make_gy(rng, rois.shape[0], 64, 14)
    a, ga, _ = run_fused(feats, rois, levels, scales, [14], 2, gys=[gy])
    b, gb, _ = run_fused(feats, rois, levels, scales, [14], 2, gys=[gy])
    assert np.array_equal(a[0], b[0])
    for x, y in zip(ga, gb):
        assert oracle.rel_err(x, y) <= BWD_TOL


@pytest.mark.parametrize("mode_name,S,sizes", [("chainer", 1, [7, 14]), ("caffe2", 2, [14]),
                                               ("caffe2", 2, [7, 14]), ("caffe2", 3, [5])])
def test_deterministic_backward_is_bit_reproducible_and_matches_oracle(mode_name, S, sizes):
    """problem.deterministic = 1: segmented reduction, no atomics.  Two runs agree bit
    for bit (also with a different CTA size), and the result matches the oracle."""
    rng, feats, rois, levels, scales = make_case(seed=21 + S, per_img=300)
    mode = _lib.COORD_CHAINER if mode_name == "chainer" else _lib.COORD_CAFFE2
    gys = [synth.make_gy(rng, rois.shape[0], feats[0].shape[1], P) for P in sizes]
    _, ga, _ = run_fused(feats, rois, levels, scales, sizes, S, mode, gys=gys, deterministic=True)
    _, gb, _ = run_fused(feats, rois, levels, scales, sizes, S, mode, gys=gys, deterministic=True)
    _, gc, _ = run_fused(feats, rois, levels, scales, sizes, S, mode, gys=gys, deterministic=True,
                         options=dict(cta_threads=256))
    _, want = oracle_fused(feats, rois, levels, scales, sizes, S, mode_name, gys)
    for a, b, c, w in zip(ga, gb, gc, want):
        assert np.array_equal(a, b) and np.array_equal(a, c)
        assert oracle.rel_err(a, w) <= BWD_TOL


def test_deterministic_backward_edge_cases():
    rng, feats, rois, levels, scales = make_case(seed=31, C=8, per_img=20)
    f = [dev(x, True) for x in feats]
    # no RoIs: zero gradients, written by the gather kernel alone
    outs, plan = _engine.forward(f, dev(rois[:0]), dev(levels[:0]), scales, [7])
    grads = _engine.backward(plan, [dev(np.zeros((0, 8, 7, 7), np.float32))], deterministic=True)
    assert all(float(g.abs().max()) == 0.0 for g in grads)
    # a RoI of another image index and clipped levels behave as in the atomic kernel
    r2 = rois.copy()
    r2[3, 0] = 9
    gy = synth.make_gy(rng, r2.shape[0], 8, 7)
    outs, plan = _engine.forward(f, dev(r2), None, scales, [7], sampling_ratio=2)
    ga = [host(g) for g in _engine.backward(plan, [dev(gy)])]
    gd = [host(g) for g in _engine.backward(plan, [dev(gy)], deterministic=True)]
    for a, d in zip(ga, gd):
        assert oracle.rel_err(d, a) <= BWD_TOL
    # pooled size 20 > 16 needs the generic path, which cannot be ordered: loud error
    outs, plan = _engine.forward(f, dev(rois), None, scales, [20])
    with pytest.raises(_lib.RpoolError):
        _engine.backward(plan, [dev(synth.make_gy(rng, rois.shape[0], 8, 20))], deterministic=True)
    assert _engine.status_flags(plan) & _lib.FLAG_DET_GENERIC
    # a caller-provided scratch (no size query, no host round trip): large enough -> same bits;
    # too small -> flagged, never a silent partial result
    outs, plan = _engine.forward(f, dev(r2), None, scales, [7], sampling_ratio=2)
    need = _engine.det_scratch_bytes(plan)
    big = torch.empty(need + 4096, dtype=torch.uint8, device="cuda")
    g1 = [host(g) for g in _engine.backward(plan, [dev(gy)], deterministic=True, det_scratch=big)]
    for a, d in zip(g1, gd):
        assert np.array_equal(a, d)
    small = torch.empty(max(need // 2, 16) // 16 * 16, dtype=torch.uint8, device="cuda")
    with pytest.raises(_lib.RpoolError):
        _engine.backward(plan, [dev(gy)], deterministic=True, det_scratch=small)
    assert _engine.status_flags(plan) & _lib.FLAG_DET_SCRATCH


@pytest.mark.parametrize("deterministic", [False, True])
def test_backward_accumulates_into_existing_gradients(deterministic):
    # rpool_problem.accumulate: the box and the mask head pooled by two separate calls add
    # into the same dense gradients (what Chainer's autograd does with the per-call results)
    rng, feats, rois, levels, scales = make_case(seed=21, C=32, per_img=80)
    gy7 = synth.make_gy(rng, rois.shape[0], 32, 7)
    gy14 = synth.make_gy(rng, rois.shape[0], 32, 14)
    f = [dev(x, True) for x in feats]
    _, plan7 = _engine.forward(f, dev(rois), None, scales, [7], sampling_ratio=2)
    _, plan14 = _engine.forward(f, dev(rois), None, scales, [14], sampling_ratio=2)
    grads = [torch.full_like(x, float("nan")) for x in f]          # the first call must overwrite
    _engine.backward(plan7, [dev(gy7, True)], out=grads, deterministic=deterministic)
    _engine.backward(plan14, [dev(gy14, True)], out=grads, accumulate=True, deterministic=deterministic)
    torch.cuda.synchronize()
    _, want = oracle_fused(feats, rois, levels, scales, [7, 14], 2, "caffe2", [gy7, gy14])
    for g, w in zip(grads, want):
        assert oracle.rel_err(host(g), w) <= BWD_TOL
    with pytest.raises(ValueError):
        _engine.backward(plan7, [dev(gy7, True)], accumulate=True)


def test_layout_conversion_kernels():
    rng = np.random.RandomState(15)
    x = rng.standard_normal((3, 37, 19, 45)).astype(np.float32)
    t = dev(x)
    cl = _engine.to_channels_last(t)
    assert cl.is_contiguous(memory_format=torch.channels_last)
    assert np.array_equal(host(cl), x)
    assert np.array_equal(host(cl.permute(0, 2, 3, 1).contiguous()), x.transpose(0, 2, 3, 1))
    back = _engine.to_nchw_contiguous(cl)
    assert back.is_contiguous() and np.array_equal(host(back), x)


# ---------------------------------------------------------------------------
# BASELINE.json sizes: every kernel configuration, both coordinate recipes
# ---------------------------------------------------------------------------
def _record(name, stats):
    """Achieved errors go to gpurun_out/parity_achieved.json (the margin to the tolerances)."""
    import json
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(path, exist_ok=True)
    path = os.path.join(path, "parity_achieved.json")
    try:
        with open(path) as f:
            d = json.load(f)
    except Exception:  # noqa: BLE001
        d = {}
    d[name] = stats
    with open(path, "w") as f:
        json.dump(d, f, indent=1, sort_keys=True)


def _full_size_case(cfg_id):
    cfg = synth.CONFIGS[cfg_id]
    rng = np.random.RandomState(cfg_id)
    feats = synth.make_pyramid(rng, cfg["n_images"], cfg["channels"], cfg["height"], cfg["width"],
                               cfg["n_levels"])
    rois = synth.make_rois(rng, cfg["n_images"], cfg["rois_per_image"], cfg["height"], cfg["width"],
                           aspect_range=cfg["aspect"])
    L = cfg["n_levels"]
    levels = oracle.levels_for_pyramid(rois[:, 1:], L)
    scales = [1.0 / s for s in synth.STRIDES[:L]]
    return cfg, rng, feats, rois, levels, scales


@pytest.mark.parametrize("cfg_id,mode_name,S", [(0, "chainer", 1), (0, "caffe2", 2),
                                                (1, "chainer", 1), (1, "caffe2", 2),
                                                (2, "chainer", 1), (2, "caffe2", 2),
                                                (3, "chainer", 1), (3, "caffe2", 2)])
def test_full_size_configs(cfg_id, mode_name, S):
    """configs[0..3] of BASELINE.json at full size against the oracle: mask head 14x14
    (fpn_roi_mask_head.py:74-78), keypoint head 14x14 on person-shaped RoIs
    (fpn_roi_keypoint_head.py:59-71,83-87), box 7x7 + mask 14x14 in one call on 16 images
    (fpn_roi_mask_head.py:57-63,74-78).  Heads are checked one at a time so that only one
    oracle result is alive at once (configs[3] pools 4 GB)."""
    cfg, rng, feats, rois, levels, scales = _full_size_case(cfg_id)
    sizes = cfg["out_sizes"]
    C, R = cfg["channels"], rois.shape[0]
    mode = _lib.COORD_CHAINER if mode_name == "chainer" else _lib.COORD_CAFFE2
    f = [dev(x, True) for x in feats]
    outs, plan = _engine.forward(f, dev(rois), None, scales, sizes, sampling_ratio=S, coord_mode=mode)
    lv, _ = _engine.read_plan(plan)
    assert np.array_equal(lv, levels)                       # level assignment: bit-exact
    assert _engine.status_flags(plan) == 0
    gys_d = [torch.rand((R, C, P, P), device="cuda", generator=torch.Generator("cuda").manual_seed(P))
             .mul_(2).sub_(1).contiguous(memory_format=torch.channels_last) for P in sizes]
    grads = [host(g) for g in _engine.backward(plan, gys_d)]
    torch.cuda.synchronize()
    del f
    shapes = [x.shape for x in feats]
    lhs = scale = 0.0
    wgrads = [np.zeros(s, np.float32) for s in shapes]
    tag = "cfg%d_%s_S%d" % (cfg_id, mode_name, S)
    for h, P in enumerate(sizes):
        o, g = host(outs[h]), host(gys_d[h])
        want = oracle.fpn_forward(feats, rois, levels, scales, P, mode_name, S, threads=oracle.max_threads())
        st = oracle.err_stats(o, want)
        _record("%s_fwd_%d" % (tag, P), st)
        assert st["max_norm"] <= FWD_TOL and st["elem_rel"] <= FWD_TOL, (P, st)
        del want
        part = oracle.fpn_backward(g, shapes, rois, levels, scales, mode_name, S,
                                   threads=oracle.max_threads())
        for l in range(len(shapes)):
            wgrads[l] += part[l]
        # adjoint identity <y, gy> == <x, gx> at full size (size-independent property)
        lhs += float((o.astype(np.float64) * g).sum())
        scale += float(np.abs(o.astype(np.float64) * g).sum())
        del o, g, part
    for l, (g, w) in enumerate(zip(grads, wgrads)):
        st = oracle.err_stats(g, w)
        _record("%s_bwd_P%d" % (tag, l + 2), st)
        assert st["max_norm"] <= BWD_TOL and st["elem_rel"] <= BWD_TOL, (l, st)
    rhs = sum(float((x.astype(np.float64) * g).sum()) for x, g in zip(feats, grads))
    assert abs(lhs - rhs) <= 2e-7 * scale


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_by_image_equals_unsharded(world):
    """BASELINE.json configs[3] dealt to 2/4/8 ranks by image (_sharding.shard_rois, image n
    -> rank n mod world): every rank's pooled rows, put back at their global row numbers,
    equal the unsharded run BIT FOR BIT, and every image's gradient matches the unsharded
    one within the backward tolerance.  The ranks run one after the other on this device;
    bench.py --shard repeats the check on real ranks."""
    from chainer_maskrcnn_b200 import _sharding
    cfg = dict(synth.CONFIGS[3], n_images=8, rois_per_image=500)        # same shapes, 8 images
    rng = np.random.RandomState(33)
    L, C, N = cfg["n_levels"], cfg["channels"], cfg["n_images"]
    shapes = synth.pyramid_shapes(N, C, cfg["height"], cfg["width"], L)
    rois = synth.make_rois(rng, N, cfg["rois_per_image"], cfg["height"], cfg["width"])
    scales = [1.0 / s for s in synth.STRIDES[:L]]
    sizes = cfg["out_sizes"]
    gen = torch.Generator("cuda").manual_seed(5)
    feats = [torch.randn(s, device="cuda", generator=gen).contiguous(memory_format=torch.channels_last)
             for s in shapes]
    gys = [torch.rand((rois.shape[0], C, P, P), device="cuda", generator=gen).mul_(2).sub_(1)
           .contiguous(memory_format=torch.channels_last) for P in sizes]
    outs, plan = _engine.forward(feats, dev(rois), None, scales, sizes, sampling_ratio=2)
    grads = _engine.backward(plan, gys)
    seen = np.zeros(rois.shape[0], np.int64)
    for rank in range(world):
        local, rows = _sharding.shard_rois(rois, N, world, rank)
        mine = _sharding.images_of_rank(N, world, rank)
        seen[rows] += 1
        idx, ridx = torch.as_tensor(mine, device="cuda"), torch.as_tensor(rows, device="cuda")
        f_loc = [f[idx].contiguous(memory_format=torch.channels_last) for f in feats]
        g_loc = [g[ridx].contiguous(memory_format=torch.channels_last) for g in gys]
        o_loc, p_loc = _engine.forward(f_loc, dev(local), None, scales, sizes, sampling_ratio=2)
        gr_loc = _engine.backward(p_loc, g_loc)
        for o, ol in zip(outs, o_loc):
            assert torch.equal(o[ridx], ol), (world, rank)                 # forward: bit for bit
        for l, (g, gl) in enumerate(zip(grads, gr_loc)):
            st = oracle.err_stats(host(gl), host(g[idx]))
            assert st["max_norm"] <= BWD_TOL and st["elem_rel"] <= BWD_TOL, (world, rank, l, st)
    assert np.all(seen == 1)                                                # every RoI on exactly one rank
