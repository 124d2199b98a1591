"""GPU: seeded random sweep of the fused call against the CPU oracle -- pyramid depth,
channel counts (including C % 4 != 0 -> generic kernel path), one or two pooled sizes
(square and rectangular), both coordinate recipes, sampling grids 0..4, given or
device-assigned levels, RoIs touching the borders.  Same tolerances as
test_gpu_parity.py (forward 1e-5, backward 1e-4, levels bit-exact)."""
import numpy as np
import pytest
import torch

import oracle
import synth
from chainer_maskrcnn_b200 import _engine, _lib

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-5
BWD_TOL = 1e-4
N_CASES = 32


def draw_case(seed):
    rng = np.random.RandomState(7000 + seed)
    L = int(rng.choice([1, 2, 3, 4, 5]))
    C = int(rng.choice([3, 4, 6, 8, 12, 20, 64, 132, 256]))
    n_img = int(rng.randint(1, 4))
    Himg, Wimg = int(rng.randint(96, 321)), int(rng.randint(96, 401))
    per_img = int(rng.randint(1, 60))
    n_heads = int(rng.randint(1, 3))
    sizes = []
    for _ in range(n_heads):
        if rng.rand() < 0.25:
            sizes.append((int(rng.randint(1, 17)), int(rng.randint(1, 17))))
        else:
            sizes.append(int(rng.choice([1, 2, 3, 5, 7, 8, 14, 16])))
    if rng.rand() < 0.4:
        mode_name, S = "chainer", 1
    else:
        mode_name, S = "caffe2", int(rng.choice([0, 1, 2, 3, 4]))
    lo = float(rng.choice([4.0, 8.0, 16.0]))
    hi = float(rng.choice([64.0, 160.0, min(Himg, Wimg) * 0.9]))
    feats = synth.make_pyramid(rng, n_img, C, Himg, Wimg, L)
    rois = synth.make_rois(rng, n_img, per_img, Himg, Wimg, size_range=(lo, max(hi, lo * 2)))
    if mode_name == "caffe2" and rois.shape[0] >= 4:
        # caffe2 semantics are defined beyond the image: push a few boxes over the borders
        k = rng.choice(rois.shape[0], size=max(1, rois.shape[0] // 8), replace=False)
        rois[k, 1:3] -= rng.uniform(0, 12, (len(k), 2)).astype(np.float32)
        rois[k, 3:5] += rng.uniform(0, 12, (len(k), 2)).astype(np.float32)
        rois[k[0], 3:5] = rois[k[0], 1:3]          # one degenerate box
    rng.shuffle(rois)
    levels = oracle.levels_for_pyramid(rois[:, 1:], L)
    given = rng.choice(["none", "i32", "f32"])
    scales = [1.0 / s for s in synth.STRIDES[:L]]
    gys = [synth.make_gy(rng, rois.shape[0], C, P) if not isinstance(P, tuple) else
           rng.uniform(-1, 1, (rois.shape[0], C) + P).astype(np.float32) for P in sizes]
    channels_last = bool(rng.rand() < 0.7)
    return dict(feats=feats, rois=rois, levels=levels, given=given, scales=scales, sizes=sizes,
                mode_name=mode_name, S=S, gys=gys, channels_last=channels_last, L=L, C=C)


def oracle_case(c):
    outs, grads = [], [np.zeros_like(f) for f in c["feats"]]
    for P, gy in zip(c["sizes"], c["gys"]):
        outs.append(oracle.fpn_forward(c["feats"], c["rois"], c["levels"], c["scales"], P,
                                       c["mode_name"], c["S"], threads=8))
        part = oracle.fpn_backward(gy, [f.shape for f in c["feats"]], c["rois"], c["levels"],
                                   c["scales"], c["mode_name"], c["S"], threads=8)
        for l in range(c["L"]):
            grads[l] += part[l]
    return outs, grads


def dev(a, channels_last):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    if channels_last and t.dim() == 4:
        t = t.contiguous(memory_format=torch.channels_last)
    return t


@pytest.mark.parametrize("seed", range(N_CASES))
def test_random_case_matches_oracle(seed):
    c = draw_case(seed)
    want, want_g = oracle_case(c)
    lv = None
    if c["given"] == "i32":
        lv = torch.from_numpy(c["levels"].astype(np.int32)).cuda()
    elif c["given"] == "f32":
        lv = torch.from_numpy(c["levels"].astype(np.float32)).cuda()
    mode = _lib.COORD_CHAINER if c["mode_name"] == "chainer" else _lib.COORD_CAFFE2
    f = [dev(x, c["channels_last"]) for x in c["feats"]]
    outs, plan = _engine.forward(f, dev(c["rois"], False), lv, c["scales"], c["sizes"],
                                 sampling_ratio=c["S"], coord_mode=mode, roi_format=_lib.ROI_YX)
    grads = _engine.backward(plan, [dev(g, c["channels_last"]) for g in c["gys"]])
    torch.cuda.synchronize()
    got_lv, _ = _engine.read_plan(plan)
    assert np.array_equal(got_lv, c["levels"])
    tag = (seed, c["L"], c["C"], c["sizes"], c["mode_name"], c["S"], c["given"])
    for o, w in zip(outs, want):
        assert tuple(o.shape) == w.shape, tag
        assert oracle.rel_err(o.cpu().numpy(), w) <= FWD_TOL, tag
    for g, w in zip(grads, want_g):
        assert oracle.rel_err(g.cpu().numpy(), w) <= BWD_TOL, tag
    # the deterministic variant, where the shapes admit it, agrees too and repeats bit for bit
    try:
        d1 = [g.cpu().numpy() for g in _engine.backward(plan, [dev(g, True) for g in c["gys"]],
                                                       deterministic=True)]
    except _lib.RpoolError as e:
        assert e.code == _lib.UNSUPPORTED, tag
        return
    d2 = [g.cpu().numpy() for g in _engine.backward(plan, [dev(g, True) for g in c["gys"]],
                                                   deterministic=True)]
    for a, b, w in zip(d1, d2, want_g):
        assert np.array_equal(a, b), tag
        assert oracle.rel_err(a, w) <= BWD_TOL, tag
