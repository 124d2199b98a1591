"""CPU: pins the oracle (oracle/) against the reference's own behaviour.

 1. bit-for-bit against the UNMODIFIED reference Python op and level mapper,
    imported from /root/reference (skipped where that tree is absent);
 2. bit-for-bit against the reference C++ compiled into oracle/_ref;
 3. bit-for-bit against the golden vectors in tests/golden/ (generated from 1+2);
 4. known-answer tests of SURVEY.md 8(c) and the reference test's properties
    (test_roi_align_2d.py: shape/dtype :39-54, numeric gradient :83-94).
"""
import os

import numpy as np
import pytest

import oracle
from oracle import reference_loader as ref
import synth

needs_reference = pytest.mark.skipif(not ref.available(), reason="/root/reference not present")
needs_ref_cpp = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _random_case(seed, N=2, C=6, H=20, W=27, R=50, scale=0.5):
    rng = np.random.RandomState(seed)
    x = rng.standard_normal((N, C, H, W)).astype(np.float32)
    rois = synth.make_rois(rng, N, R // N, int(H / scale), int(W / scale), size_range=(1.0, 40.0))
    return x, oracle.roi_yx_to_xy(rois), scale


# ---- 1. unmodified reference ------------------------------------------------
@needs_reference
@pytest.mark.parametrize("seed,outh,outw", [(0, 7, 7), (1, 5, 7), (2, 14, 14), (3, 1, 1)])
def test_chainer_path_equals_reference_bitwise(seed, outh, outw):
    x, rois, scale = _random_case(seed)
    rois = np.concatenate([rois, np.array([[0, 3, 3, 3, 3], [1, 2.5, 4, 2.6, 9]], np.float32)])
    y_ref = ref.reference_forward(x, rois, outh, outw, scale)
    y = oracle.forward_chainer(x, rois, outh, outw, scale)
    assert y.dtype == np.float32 and np.array_equal(y, y_ref)
    gy = np.random.RandomState(seed + 100).uniform(-1, 1, y.shape).astype(np.float32)
    g_ref = ref.reference_backward(gy, x, rois, outh, outw, scale)
    g = oracle.backward_chainer(gy, rois, x.shape, scale)
    assert np.array_equal(g, g_ref)
    assert np.array_equal(oracle.backward_chainer(gy, rois, x.shape, scale, threads=3), g_ref)
    assert np.array_equal(oracle.forward_chainer(x, rois, outh, outw, scale, threads=3), y_ref)


@needs_reference
def test_reference_raises_where_oracle_raises():
    x = np.zeros((1, 2, 8, 8), np.float32)
    rois = np.array([[0, 0, 0, 20, 20]], np.float32)   # reaches past the map
    with pytest.raises(IndexError):
        ref.reference_forward(x, rois, 4, 4, 1.0)
    with pytest.raises(IndexError):
        oracle.forward_chainer(x, rois, 4, 4, 1.0)


@needs_reference
def test_level_mapper_equals_reference():
    rng = np.random.RandomState(3)
    b = np.zeros((20000, 4), np.float32)
    b[:, :2] = rng.uniform(0, 1000, (20000, 2))
    b[:, 2:] = b[:, :2] + np.exp(rng.uniform(np.log(1), np.log(1200), (20000, 2)))
    f = ref.reference_level_mapper()
    assert np.array_equal(f(b), oracle.map_rois_to_fpn_levels(b))
    assert np.array_equal(f(b, 1, 3), oracle.map_rois_to_fpn_levels(b, 1, 3))


# ---- 2. reference C++ --------------------------------------------------------
@needs_ref_cpp
@pytest.mark.parametrize("S", [1, 2, 3, 4, 0])
@pytest.mark.parametrize("outh,outw", [(7, 7), (5, 7), (14, 14)])
def test_caffe2_restatement_equals_reference_cpp_bitwise(S, outh, outw):
    x, rois, scale = _random_case(10 + S)
    # include boxes hanging over every border: caffe2 semantics define them
    extra = np.array([[0, -8, -8, 10, 10], [1, 40, 30, 70, 50], [0, -30, 5, -20, 9],
                      [1, 10, 10, 10, 10]], np.float32)
    rois = np.concatenate([rois, extra])
    a = oracle.forward_caffe2(x, rois, outh, outw, scale, S)
    b = oracle.ref_caffe2_forward(x, rois, outh, outw, scale, S)
    assert np.array_equal(a, b)
    assert np.array_equal(oracle.forward_caffe2(x, rois, outh, outw, scale, S, threads=4), b)


# ---- 2b. randomised sweeps against the reference itself ----------------------
def _sweep_case(seed):
    rng = np.random.RandomState(9000 + seed)
    N, C = int(rng.randint(1, 4)), int(rng.randint(1, 9))
    H, W = int(rng.randint(4, 40)), int(rng.randint(4, 40))
    scale = float(rng.choice([1.0, 0.5, 0.25, 0.125, 0.6, 1.0 / 3.0]))
    outh, outw = int(rng.randint(1, 16)), int(rng.randint(1, 16))
    R = int(rng.randint(1, 40))
    x = rng.standard_normal((N, C, H, W)).astype(np.float32)
    # boxes inside the map (the parity domain of the NumPy path), some of them degenerate
    b = rng.randint(0, N, R).astype(np.float32)
    x1 = rng.uniform(0, (W - 1) / scale, R)
    y1 = rng.uniform(0, (H - 1) / scale, R)
    x2 = x1 + rng.uniform(0, 1, R) * ((W - 1) / scale - x1)
    y2 = y1 + rng.uniform(0, 1, R) * ((H - 1) / scale - y1)
    deg = rng.rand(R) < 0.15
    x2[deg] = x1[deg]
    rois = np.stack([b, x1, y1, x2, y2], 1).astype(np.float32)
    return rng, x, rois, outh, outw, scale


@needs_reference
@pytest.mark.parametrize("seed", range(24))
def test_random_sweep_chainer_path_equals_reference_bitwise(seed):
    rng, x, rois, outh, outw, scale = _sweep_case(seed)
    try:
        y_ref = ref.reference_forward(x, rois, outh, outw, scale)
    except IndexError:
        with pytest.raises(IndexError):
            oracle.forward_chainer(x, rois, outh, outw, scale)
        return
    y = oracle.forward_chainer(x, rois, outh, outw, scale)
    assert np.array_equal(y, y_ref)
    gy = rng.uniform(-1, 1, y.shape).astype(np.float32)
    assert np.array_equal(oracle.backward_chainer(gy, rois, x.shape, scale),
                          ref.reference_backward(gy, x, rois, outh, outw, scale))


@needs_ref_cpp
@pytest.mark.parametrize("seed", range(24))
def test_random_sweep_caffe2_restatement_equals_reference_cpp_bitwise(seed):
    rng, x, rois, outh, outw, scale = _sweep_case(seed)
    # caffe2 semantics are defined beyond the map: push some boxes over the borders
    k = rng.rand(rois.shape[0]) < 0.3
    rois[k, 1:3] -= rng.uniform(0, 20, (int(k.sum()), 2)).astype(np.float32)
    rois[k, 3:5] += rng.uniform(0, 20, (int(k.sum()), 2)).astype(np.float32)
    S = int(rng.choice([0, 1, 2, 3, 4]))
    a = oracle.forward_caffe2(x, rois, outh, outw, scale, S)
    assert np.array_equal(a, oracle.ref_caffe2_forward(x, rois, outh, outw, scale, S))
    # its backward is defined as the adjoint of that forward: <y, gy> == <x, gx>
    gy = rng.uniform(-1, 1, a.shape).astype(np.float32)
    gx = oracle.backward_caffe2(gy, rois, x.shape, scale, S)
    lhs = float(np.sum(a.astype(np.float64) * gy))
    rhs = float(np.sum(x.astype(np.float64) * gx))
    assert abs(lhs - rhs) <= 1e-3 * max(1.0, abs(lhs))


# ---- 3. golden vectors -------------------------------------------------------
def test_golden_reference_fixture(golden_dir):
    d = _load(golden_dir, "reference_fixture.npz")
    outh, outw, scale = int(d["outh"]), int(d["outw"]), float(d["scale"])
    y = oracle.forward_chainer(d["x"], d["rois"], outh, outw, scale)
    assert y.shape == d["gy"].shape and y.dtype == np.float32   # test_roi_align_2d.py:49-52
    assert np.array_equal(y, d["y"])
    assert np.array_equal(oracle.backward_chainer(d["gy"], d["rois"], d["x"].shape, scale), d["gx"])
    for S in (1, 2, 3):
        assert np.array_equal(oracle.forward_caffe2(d["x"], d["rois"], outh, outw, scale, S),
                              d["y_caffe2_s%d" % S])
    # NumPy path and caffe2 path agree to cross-implementation noise
    assert oracle.rel_err(d["y"], d["y_caffe2_s1"]) < 1e-5


def test_golden_fpn_small(golden_dir):
    d = _load(golden_dir, "fpn_small.npz")
    feats = [d["feat%d" % l] for l in range(4)]
    scales = [float(s) for s in d["scales"]]
    lv = oracle.levels_for_pyramid(d["rois"][:, 1:], 4)
    assert np.array_equal(lv, d["levels"])
    assert np.array_equal(oracle.map_rois_to_fpn_levels(d["rois"][:, 1:]), d["levels_f32"])
    for P in (7, 14):
        y = oracle.fpn_forward(feats, d["rois"], lv, scales, P)
        assert np.array_equal(y, d["y%d" % P])
        gx = oracle.fpn_backward(d["gy%d" % P], [f.shape for f in feats], d["rois"], lv, scales)
        for l in range(4):
            assert np.array_equal(gx[l], d["gx%d_l%d" % (P, l)])
        y2 = oracle.fpn_forward(feats, d["rois"], lv, scales, P, mode="caffe2", sampling_ratio=2)
        assert np.array_equal(y2, d["y%d_caffe2_s2" % P])


def test_golden_sweep(golden_dir):
    # random small cases generated from the reference (tests/golden/make_golden.py: sweep)
    d = _load(golden_dir, "sweep.npz")
    for i in range(int(d["n_cases"])):
        g = lambda k: d["c%d_%s" % (i, k)]
        outh, outw, S = (int(v) for v in g("geom"))
        scale = float(g("scale"))
        assert np.array_equal(oracle.forward_chainer(g("x"), g("rois"), outh, outw, scale), g("y")), i
        assert np.array_equal(oracle.backward_chainer(g("gy"), g("rois"), g("x").shape, scale), g("gx")), i
        assert np.array_equal(oracle.forward_caffe2(g("x"), g("rois_c2"), outh, outw, scale, S), g("y_c2")), i


def test_golden_levels(golden_dir):
    d = _load(golden_dir, "levels.npz")
    assert np.array_equal(oracle.map_rois_to_fpn_levels(d["boxes"]), d["levels"])
    assert np.array_equal(oracle.map_rois_to_fpn_levels(d["boxes"], 0, 3), d["levels_kmax3"])
    thr = oracle.level_area_thresholds()
    assert np.array_equal(thr, d["thresholds"])
    assert [hex(int(t.view(np.uint32))) for t in thr] == \
        ["0x4443ff31", "0x4543ff98", "0x4643ffc9", "0x4743ffe4"]
    assert np.all(np.diff(thr) > 0)
    # the threshold form reproduces the mapper exactly on the golden boxes
    b = d["boxes"]
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    with np.errstate(invalid="ignore"):
        lv = (area[:, None] >= thr[None, :]).sum(1).astype(np.float32)
    assert np.array_equal(lv, d["levels"])


# ---- 4. known answers and properties ----------------------------------------
def test_known_answers():
    x = (10 * np.arange(6)[:, None] + np.arange(6)[None, :]).astype(np.float32)[None, None]
    kat = [([0, 0, 0, 4, 4], [11, 13, 31, 33]),
           ([0, .5, .5, 3.5, 3.5], [13.75, 15.25, 28.75, 30.25]),
           ([0, 1, 2, 1, 2], [23.75, 24.25, 28.75, 29.25]),
           ([0, 2, 0, 5, 5], [15.25, 16.75, 40.25, 41.75])]
    for roi, want in kat:
        y = oracle.forward_chainer(x, np.array([roi], np.float32), 2, 2, 1.0)
        assert np.array_equal(y.ravel(), np.array(want, np.float32)), roi
    yx = np.array([[0, 4, 8, 20, 24]], np.float32)
    y = oracle.forward_chainer(x, oracle.roi_yx_to_xy(yx), 2, 2, 0.25)
    assert np.array_equal(y.ravel(), np.array([23, 25, 43, 45], np.float32))
    g = oracle.backward_chainer(np.array([[[[1, 2], [3, 4]]]], np.float32),
                                np.array([[0, .5, .5, 3.5, 3.5]], np.float32), x.shape, 1.0)[0, 0]
    want = np.zeros((6, 6), np.float32)
    want[1:4, 1:4] = [[.5625, .5625, 1.125], [.75, .625, 1.125], [1.6875, 1.3125, 2.25]]
    assert np.array_equal(g, want)


@pytest.mark.parametrize("mode,S", [("chainer", 1), ("caffe2", 1), ("caffe2", 2)])
def test_linear_field_is_reproduced(mode, S):
    H, W = 30, 41
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    x = (3.0 * yy + 0.5 * xx).astype(np.float32)[None, None]
    rng = np.random.RandomState(0)
    rois = oracle.roi_yx_to_xy(synth.make_rois(rng, 1, 64, H - 1, W - 1, size_range=(2.0, 25.0)))
    P = 6
    if mode == "chainer":
        y = oracle.forward_chainer(x, rois, P, P, 1.0)
    else:
        y = oracle.forward_caffe2(x, rois, P, P, 1.0, S)
    ys, xs = rois[:, 2], rois[:, 1]
    rh = np.maximum(rois[:, 4] - ys, 1)
    rw = np.maximum(rois[:, 3] - xs, 1)
    cy = ys[:, None] + (np.arange(P)[None, :] + 0.5) * rh[:, None] / P
    cx = xs[:, None] + (np.arange(P)[None, :] + 0.5) * rw[:, None] / P
    want = 3.0 * cy[:, :, None] + 0.5 * cx[:, None, :]
    assert np.abs(y[:, 0] - want).max() < 2e-4


@pytest.mark.parametrize("mode,S", [("chainer", 1), ("caffe2", 2)])
def test_backward_is_adjoint_and_matches_numeric_gradient(mode, S):
    x, rois, scale = _random_case(42, N=2, C=3, H=12, W=9, R=20)
    outh, outw = 5, 7
    if mode == "chainer":
        f = lambda v: oracle.forward_chainer(v, rois, outh, outw, scale)
        b = lambda g: oracle.backward_chainer(g, rois, x.shape, scale)
    else:
        f = lambda v: oracle.forward_caffe2(v, rois, outh, outw, scale, S)
        b = lambda g: oracle.backward_caffe2(g, rois, x.shape, scale, S)
    y = f(x)
    gy = np.random.RandomState(1).uniform(-1, 1, y.shape).astype(np.float32)
    gx = b(gy)
    lhs = float((y.astype(np.float64) * gy).sum())
    rhs = float((x.astype(np.float64) * gx).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))
    # numeric gradient at the reference's tolerances (test_roi_align_2d.py:37)
    rng = np.random.RandomState(2)
    eps = 1e-2
    for _ in range(25):
        idx = tuple(rng.randint(s) for s in x.shape)
        xp, xm = x.copy(), x.copy()
        xp[idx] += eps
        xm[idx] -= eps
        num = float(((f(xp).astype(np.float64) - f(xm)) * gy).sum() / (2 * eps))
        assert abs(num - gx[idx]) <= 1e-3 + 1e-2 * abs(num)


def test_empty_and_threads():
    x = np.ones((1, 2, 5, 5), np.float32)
    rois = np.zeros((0, 5), np.float32)
    assert oracle.forward_chainer(x, rois, 3, 3, 1.0).shape == (0, 2, 3, 3)
    assert np.all(oracle.backward_chainer(np.zeros((0, 2, 3, 3), np.float32), rois, x.shape, 1.0) == 0)
    assert oracle.max_threads() >= 1
