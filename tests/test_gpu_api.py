"""GPU: the reference-shaped Python API (functions/, model/) on top of the C ABI."""
import numpy as np
import pytest
import torch

import oracle
import synth
import chainer_maskrcnn_b200 as pkg
from chainer_maskrcnn_b200 import _lib
from chainer_maskrcnn_b200.functions.roi_align.roi_align_2d import ROIAlign2D, roi_align_2d
from chainer_maskrcnn_b200.functions.roi_align_2d_yx import _roi_align_2d_yx

pytestmark = pytest.mark.gpu


def _fixture(seed=0, N=3, C=32, H=12, W=8):
    # the reference test's fixture (test_roi_align_2d.py:17-37) with a fixed seed
    rng = np.random.RandomState(seed)
    x = np.arange(N * C * H * W, dtype=np.float32).reshape(N, C, H, W)
    rng.shuffle(x)
    x = (2 * x / x.size - 1).astype(np.float32)
    rois = np.tile(np.array([[0, 1, 1, 6, 6], [2, 6, 2, 7, 11], [1, 3, 1, 5, 10], [0, 3, 3, 3, 3]],
                            np.float32), (15, 1))
    gy = rng.uniform(-1, 1, (rois.shape[0], C, 5, 7)).astype(np.float32)
    return x, rois, gy, 5, 7, 0.6


def test_forward_shape_dtype_and_tuple_conventions():
    x, rois, gy, outh, outw, scale = _fixture()
    f = ROIAlign2D(outh, outw, scale)
    out = f.forward_gpu((torch.from_numpy(x).cuda(), torch.from_numpy(rois).cuda()))
    assert isinstance(out, tuple) and len(out) == 1                 # roi_align_2d.py:146
    assert out[0].dtype == torch.float32 and tuple(out[0].shape) == gy.shape
    assert f._bottom_data_shape == x.shape                           # roi_align_2d.py:94
    back = f.backward_gpu((None, torch.from_numpy(rois).cuda()), (torch.from_numpy(gy).cuda(),))
    assert isinstance(back, tuple) and len(back) == 2 and back[1] is None   # :281
    assert tuple(back[0].shape) == x.shape
    assert oracle.rel_err(out[0].cpu().numpy(), oracle.forward_chainer(x, rois, outh, outw, scale)) <= 1e-5
    assert oracle.rel_err(back[0].cpu().numpy(),
                          oracle.backward_chainer(gy, rois, x.shape, scale)) <= 1e-4


def test_autograd_matches_oracle_and_numeric_gradient():
    x, rois, gy, outh, outw, scale = _fixture(C=4)
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    y = roi_align_2d(xt, torch.from_numpy(rois).cuda(), outh, outw, scale)
    y.backward(torch.from_numpy(gy).cuda())
    g = xt.grad.cpu().numpy()
    assert oracle.rel_err(g, oracle.backward_chainer(gy, rois, x.shape, scale)) <= 1e-4
    # numeric gradient on the device op at the reference's tolerances (:37)
    rng = np.random.RandomState(1)
    eps = 1e-2
    r = torch.from_numpy(rois).cuda()
    gyt = torch.from_numpy(gy).cuda()
    for _ in range(10):
        idx = tuple(rng.randint(s) for s in x.shape)
        xp, xm = x.copy(), x.copy()
        xp[idx] += eps
        xm[idx] -= eps
        fp = roi_align_2d(torch.from_numpy(xp).cuda(), r, outh, outw, scale).double()
        fm = roi_align_2d(torch.from_numpy(xm).cuda(), r, outh, outw, scale).double()
        num = float(((fp - fm) * gyt).sum() / (2 * eps))
        assert abs(num - g[idx]) <= 1e-3 + 1e-2 * abs(num)


def test_host_arrays_roundtrip_through_gpu():
    x, rois, gy, outh, outw, scale = _fixture()
    f = ROIAlign2D(outh, outw, scale)
    (y,) = f.forward_cpu((x, rois))
    assert isinstance(y, np.ndarray) and y.dtype == np.float32 and y.shape == gy.shape
    assert oracle.rel_err(y, oracle.forward_chainer(x, rois, outh, outw, scale)) <= 1e-5
    gx, none = f.backward_cpu((x, rois), (gy,))
    assert none is None and isinstance(gx, np.ndarray) and gx.shape == x.shape
    assert oracle.rel_err(gx, oracle.backward_chainer(gy, rois, x.shape, scale)) <= 1e-4
    y2 = roi_align_2d(x, rois, outh, outw, scale)
    assert np.array_equal(np.asarray(y2), np.asarray(y))
    (y3,) = f.forward((x, rois))
    assert np.array_equal(np.asarray(y3), np.asarray(y))


def test_forward_cpu2_is_the_cpp_port_entry():
    # roi_align_2d.py:34-37: caffe2 semantics, sampling_ratio 1 (caffe2_roi_align.cpp:240)
    x, rois, gy, outh, outw, scale = _fixture(seed=2)
    f = ROIAlign2D(outh, outw, scale, sampling_ratio=2)        # its own ratio is not used by this entry
    y, = f.forward_cpu2((x, rois))
    assert isinstance(y, np.ndarray) and y.shape == (rois.shape[0], x.shape[1], outh, outw)
    want = oracle.forward_caffe2(x, rois, outh, outw, scale, 1)
    assert oracle.rel_err(y, want) <= 1e-5
    if oracle.have_ref():                                      # the reference's own compiled C++
        assert oracle.rel_err(y, oracle.ref_caffe2_forward(x, rois, outh, outw, scale, 1)) <= 1e-5


def test_yx_wrapper_equals_permuted_call():
    x, rois, gy, outh, outw, scale = _fixture()
    yx = rois[:, [0, 2, 1, 4, 3]].copy()
    a = _roi_align_2d_yx(torch.from_numpy(x).cuda(), torch.from_numpy(yx).cuda(), outh, outw, scale)
    b = roi_align_2d(torch.from_numpy(x).cuda(), torch.from_numpy(rois).cuda(), outh, outw, scale)
    assert torch.equal(a, b)
    # the SURVEY 8(c) known answer for the wrapper
    field = (10 * np.arange(6)[:, None] + np.arange(6)[None, :]).astype(np.float32)[None, None]
    out = _roi_align_2d_yx(torch.from_numpy(field).cuda(),
                           torch.tensor([[0, 4, 8, 20, 24]], dtype=torch.float32).cuda(), 2, 2, 0.25)
    assert out.flatten().tolist() == [23.0, 25.0, 43.0, 45.0]


def test_fused_head_call_replaces_the_per_roi_loop():
    rng = np.random.RandomState(3)
    n_img, C, H, W, L = 2, 32, 160, 224, 5                 # five levels: p2..p6
    feats = synth.make_pyramid(rng, n_img, C, H, W, L)
    rois = synth.make_rois(rng, n_img, 50, H, W, size_range=(8.0, 300.0))
    from chainer_maskrcnn_b200.model.extractor.feature_pyramid_network import spatial_scales as scales
    ft = [torch.from_numpy(f).cuda().requires_grad_(True) for f in feats]
    rt = torch.from_numpy(rois).cuda()
    levels_f = pkg.map_rois_to_fpn_levels(rt[:, 1:])        # float32, like the reference
    assert levels_f.dtype == torch.float32
    want_lv = oracle.map_rois_to_fpn_levels(rois[:, 1:])
    assert np.array_equal(levels_f.cpu().numpy(), want_lv)
    pooler = pkg.FPNRoIPooling(7, 14)
    box, mask = pooler(ft, rt, levels_f, scales, train=True)
    lv = np.clip(want_lv, 0, L - 1).astype(np.int32)
    assert oracle.rel_err(box.detach().cpu().numpy(), oracle.fpn_forward(feats, rois, lv, scales, 7)) <= 1e-5
    assert oracle.rel_err(mask.detach().cpu().numpy(), oracle.fpn_forward(feats, rois, lv, scales, 14)) <= 1e-5
    gb = synth.make_gy(rng, rois.shape[0], C, 7)
    gm = synth.make_gy(rng, rois.shape[0], C, 14)
    (box * torch.from_numpy(gb).cuda()).sum().backward(retain_graph=True)
    (mask * torch.from_numpy(gm).cuda()).sum().backward()
    shapes = [f.shape for f in feats]
    w7 = oracle.fpn_backward(gb, shapes, rois, lv, scales)
    w14 = oracle.fpn_backward(gm, shapes, rois, lv, scales)
    for l in range(L):
        assert oracle.rel_err(ft[l].grad.cpu().numpy(), w7[l] + w14[l]) <= 1e-4
    # test mode: box only, features cached for predict_mask (fpn_roi_mask_head.py:85-95)
    box2 = pooler(ft, rt, levels_f, scales, train=False)
    assert torch.equal(box2, box)
    mask2 = pooler.predict_mask(levels_f, rt, scales)
    assert torch.equal(mask2, mask)
    # levels=None: assigned on the device, same result
    box3 = pkg.fpn_roi_align(ft, rt, None, scales, 7)
    assert torch.equal(box3, box)


def test_keypoint_head_dispatch_including_the_single_level_shortcut():
    # fpn_roi_keypoint_head.py:59-71: when every RoI has the same level the box
    # features come from ONE batched call on x[0] / spatial_scales[0], whatever the
    # level is; the mask branch (:83-87) always follows the per-RoI level
    rng = np.random.RandomState(5)
    n_img, C, H, W, L = 2, 32, 160, 224, 4
    feats = synth.make_pyramid(rng, n_img, C, H, W, L)
    from chainer_maskrcnn_b200.model.extractor.feature_pyramid_network import spatial_scales as scales
    ft = [torch.from_numpy(f).cuda() for f in feats]
    # person-shaped boxes that all land on one level, and not the finest
    rois = synth.make_rois(rng, n_img, 40, H, W, size_range=(60.0, 100.0), aspect_range=(1.5, 3.0))
    lv = oracle.levels_for_pyramid(rois[:, 1:], L)
    assert len(np.unique(lv)) == 1 and lv[0] == 2
    rt = torch.from_numpy(rois).cuda()
    lvt = torch.from_numpy(lv.astype(np.float32)).cuda()
    pooler = pkg.FPNRoIKeypointPooling(7, 14)
    box, mask = pooler(ft, rt, lvt, scales, train=True)
    zeros = np.zeros_like(lv)
    assert oracle.rel_err(box.cpu().numpy(), oracle.fpn_forward(feats, rois, zeros, scales, 7)) <= 1e-5
    assert oracle.rel_err(mask.cpu().numpy(), oracle.fpn_forward(feats, rois, lv, scales, 14)) <= 1e-5
    # without the shortcut: box by level, like the mask head
    box_l, mask_l = pkg.FPNRoIKeypointPooling(7, 14, reference_quirk=False)(ft, rt, lvt, scales)
    assert oracle.rel_err(box_l.cpu().numpy(), oracle.fpn_forward(feats, rois, lv, scales, 7)) <= 1e-5
    assert torch.equal(mask_l, mask)
    # mixed levels: the per-RoI loop (:66-71), row r <-> RoI r
    rois2 = synth.make_rois(rng, n_img, 40, H, W, size_range=(8.0, 300.0), aspect_range=(1.5, 3.0))
    lv2 = oracle.levels_for_pyramid(rois2[:, 1:], L)
    assert len(np.unique(lv2)) > 1
    rt2 = torch.from_numpy(rois2).cuda()
    lvt2 = torch.from_numpy(lv2.astype(np.float32)).cuda()
    box2, mask2 = pooler(ft, rt2, lvt2, scales, train=True)
    assert oracle.rel_err(box2.cpu().numpy(), oracle.fpn_forward(feats, rois2, lv2, scales, 7)) <= 1e-5
    assert oracle.rel_err(mask2.cpu().numpy(), oracle.fpn_forward(feats, rois2, lv2, scales, 14)) <= 1e-5
    # test mode + predict_mask (:93-104)
    b3 = pooler(ft, rt, lvt, scales, train=False)
    assert torch.equal(b3, box)
    assert torch.equal(pooler.predict_mask(lvt, rt, scales), mask)


@pytest.mark.parametrize("C,P,scale", [(490, 7, 1 / 16.), (1024, 7, 1 / 16.), (1024, 14, 1 / 16.)])
def test_single_level_heads_thin_and_wide_maps(C, P, scale):
    # light_roi_mask_head.py:26,87-92,116-117 (thin 490-channel map) and
    # resnet_roi_mask_head.py:61-62 (1024-channel C4 map): one batched call, one map
    rng = np.random.RandomState(C + P)
    N, H, W = 2, 38, 50
    x = rng.randn(N, C, H, W).astype(np.float32)
    rois = synth.make_rois(rng, N, 24, H * 16, W * 16, size_range=(32.0, 400.0))
    gy = synth.make_gy(rng, rois.shape[0], C, P)
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    y = _roi_align_2d_yx(xt, torch.from_numpy(rois).cuda(), P, P, scale)
    assert tuple(y.shape) == (rois.shape[0], C, P, P)
    rois_xy = rois[:, [0, 2, 1, 4, 3]]
    assert oracle.rel_err(y.detach().cpu().numpy(), oracle.forward_chainer(x, rois_xy, P, P, scale, threads=8)) <= 1e-5
    (y * torch.from_numpy(gy).cuda()).sum().backward()
    assert oracle.rel_err(xt.grad.cpu().numpy(), oracle.backward_chainer(gy, rois_xy, x.shape, scale, threads=8)) <= 1e-4


def test_step_is_cuda_graph_capturable_and_replays_on_new_rois():
    # plan + forward + zero fill + backward have no host synchronisation and no host-side
    # data-dependent sizes: capture once, replay after the RoIs changed in place
    rng = np.random.RandomState(11)
    n_img, C, H, W, L = 2, 32, 160, 224, 4
    feats = synth.make_pyramid(rng, n_img, C, H, W, L)
    scales = [1.0 / s for s in synth.STRIDES[:L]]
    rois_a = synth.make_rois(rng, n_img, 60, H, W, size_range=(8.0, 300.0))
    rois_b = synth.make_rois(rng, n_img, 60, H, W, size_range=(8.0, 300.0))
    gy = synth.make_gy(rng, rois_a.shape[0], C, 7)
    from chainer_maskrcnn_b200 import _engine
    ft = [torch.from_numpy(f).cuda().contiguous(memory_format=torch.channels_last) for f in feats]
    rois = torch.from_numpy(rois_a).cuda()
    gyt = torch.from_numpy(gy).cuda().contiguous(memory_format=torch.channels_last)
    grads = [torch.empty_like(f) for f in ft]

    def step():
        outs, plan = _engine.forward(ft, rois, None, scales, [7], sampling_ratio=2)
        _engine.backward(plan, [gyt], out=grads)
        return outs[0]

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = step()
    for r in (rois_b, rois_a, rois_b):
        rois.copy_(torch.from_numpy(r).cuda())
        g.replay()
        torch.cuda.synchronize()
        lv = oracle.levels_for_pyramid(r[:, 1:], L)
        want = oracle.fpn_forward(feats, r, lv, scales, 7, "caffe2", 2)
        assert oracle.rel_err(out.cpu().numpy(), want) <= 1e-5
        want_g = oracle.fpn_backward(gy, [f.shape for f in feats], r, lv, scales, "caffe2", 2)
        for l in range(L):
            assert oracle.rel_err(grads[l].cpu().numpy(), want_g[l]) <= 1e-4


@pytest.mark.parametrize("fork", [False, True])
@pytest.mark.parametrize("graph", [True, False])
@pytest.mark.parametrize("sizes,S", [([7], 2), ([7, 14], 1)])
def test_fused_step_helper_forks_the_zero_fill_and_replays(graph, sizes, S, fork):
    """pkg.FusedStep: static buffers, gradient zero fill on a side stream (rpool_zero_fill +
    accumulate), the step captured into a CUDA graph; new RoIs / gradients are written into
    the same tensors between replays."""
    rng = np.random.RandomState(12)
    n_img, C, H, W, L = 2, 32, 160, 224, 4
    feats = synth.make_pyramid(rng, n_img, C, H, W, L)
    scales = [1.0 / s for s in synth.STRIDES[:L]]
    rois_a = synth.make_rois(rng, n_img, 70, H, W, size_range=(8.0, 300.0))
    rois_b = synth.make_rois(rng, n_img, 70, H, W, size_range=(8.0, 300.0))
    gys = [synth.make_gy(rng, rois_a.shape[0], C, P) for P in sizes]
    cl = lambda a: torch.from_numpy(a).cuda().contiguous(memory_format=torch.channels_last)
    rois = torch.from_numpy(rois_a).cuda()
    step = pkg.FusedStep([cl(f) for f in feats], rois, None, scales, sizes, S, gys=[cl(g) for g in gys],
                         graph=graph, fork_zero_fill=fork)
    assert (step.graph is not None) == graph and step.fork == fork
    mode = "chainer" if S == 1 else "caffe2"
    for r in (rois_a, rois_b, rois_a):
        rois.copy_(torch.from_numpy(r).cuda())
        for g in step.grads:
            g.fill_(float("nan"))                 # the forked fill must overwrite everything
        n0 = _lib.launch_count()
        outs, grads = step.run()
        torch.cuda.synchronize()
        if not graph:
            # plan (one launch up to 1024 RoIs) + forward + zero fill + one backward launch per pooled size
            assert _lib.launch_count() - n0 == 3 + len(sizes)
        lv = oracle.levels_for_pyramid(r[:, 1:], L)
        want_g = [np.zeros_like(f) for f in feats]
        for o, P, gy in zip(outs, sizes, gys):
            assert oracle.rel_err(o.cpu().numpy(), oracle.fpn_forward(feats, r, lv, scales, P, mode, S)) <= 1e-5
            for l, part in enumerate(oracle.fpn_backward(gy, [f.shape for f in feats], r, lv, scales, mode, S)):
                want_g[l] += part
        for g, w in zip(grads, want_g):
            assert oracle.rel_err(g.cpu().numpy(), w) <= 1e-4


def test_backward_gpu_replans_when_the_rois_changed_in_place():
    """the forward plan is reused in backward_gpu only for the same, unmodified
    RoI tensor (storage, shape and version counter)."""
    x, rois, gy, outh, outw, scale = _fixture(seed=3)
    f = ROIAlign2D(outh, outw, scale)
    xt, rt = torch.from_numpy(x).cuda(), torch.from_numpy(rois).cuda()
    f.forward_gpu((xt, rt))
    moved = rois.copy()
    moved[:, 1:] = moved[:, 1:] * 0.5 + 0.5
    rt.copy_(torch.from_numpy(moved).cuda())              # in place: same data_ptr
    gx, _ = f.backward_gpu((None, rt), (torch.from_numpy(gy).cuda(),))
    want = oracle.backward_chainer(gy, moved, x.shape, scale)
    assert oracle.rel_err(gx.cpu().numpy(), want) <= 1e-4
    gx2, _ = f.backward_gpu((None, rt[:5]), (torch.from_numpy(gy[:5]).cuda(),))     # a view: other shape
    assert oracle.rel_err(gx2.cpu().numpy(), oracle.backward_chainer(gy[:5], moved[:5], x.shape, scale)) <= 1e-4


def test_host_fused_call_and_launch_counter():
    rng = np.random.RandomState(4)
    feats = synth.make_pyramid(rng, 1, 16, 128, 128, 4)
    rois = synth.make_rois(rng, 1, 40, 128, 128, size_range=(8.0, 120.0))
    scales = [1.0 / s for s in synth.STRIDES[:4]]
    gys = [synth.make_gy(rng, 40, 16, 7)]
    n0 = _lib.launch_count()
    pooled, grads = pkg.fpn_roi_align_host(feats, rois, None, scales, [7], 2, gys=gys)
    assert _lib.launch_count() - n0 >= 4 + 3          # 4 layout conversions + plan + fwd + zero + bwd
    lv = oracle.levels_for_pyramid(rois[:, 1:], 4)
    assert oracle.rel_err(pooled[0], oracle.fpn_forward(feats, rois, lv, scales, 7, "caffe2", 2)) <= 1e-5
    want = oracle.fpn_backward(gys[0], [f.shape for f in feats], rois, lv, scales, "caffe2", 2)
    for g, w in zip(grads, want):
        assert g.shape == w.shape and oracle.rel_err(g, w) <= 1e-4


@pytest.mark.parametrize("order", ["image_major", "interleaved"])
def test_host_call_pipelined_over_image_groups(order):
    # image-major RoIs: the host call is pipelined over groups of images; any other order
    # takes the single-group path.  Same results either way, rows in input order.
    rng = np.random.RandomState(6)
    n_img, C, H, W, L = 5, 16, 128, 160, 3
    feats = synth.make_pyramid(rng, n_img, C, H, W, L)
    rois = synth.make_rois(rng, n_img, 30, H, W, size_range=(8.0, 150.0))
    rois = rois[rois[:, 0] != 3]                      # an image without RoIs
    if order == "interleaved":
        rng.shuffle(rois)
    from chainer_maskrcnn_b200.functions.fpn_roi_align import _image_groups
    assert len(_image_groups(rois, n_img)) == (4 if order == "image_major" else 1)
    scales = [1.0 / s for s in synth.STRIDES[:L]]
    lv = oracle.levels_for_pyramid(rois[:, 1:], L)
    gys = [synth.make_gy(rng, rois.shape[0], C, 7), synth.make_gy(rng, rois.shape[0], C, 14)]
    for levels in (None, lv.astype(np.float32)):
        pooled, grads = pkg.fpn_roi_align_host(feats, rois, levels, scales, [7, 14], 2, gys=gys)
        want_g = [np.zeros_like(f) for f in feats]
        for P, o, gy in zip((7, 14), pooled, gys):
            assert o.shape == (rois.shape[0], C, P, P)
            assert oracle.rel_err(o, oracle.fpn_forward(feats, rois, lv, scales, P, "caffe2", 2)) <= 1e-5
            for l, part in enumerate(oracle.fpn_backward(gy, [f.shape for f in feats], rois, lv, scales,
                                                         "caffe2", 2)):
                want_g[l] += part
        for g, w in zip(grads, want_g):
            assert g.shape == w.shape and oracle.rel_err(g, w) <= 1e-4
            assert np.all(g[3] == 0)
    # forward only
    pooled, grads = pkg.fpn_roi_align_host(feats, rois, None, scales, 7, 1)
    assert grads is None
    assert oracle.rel_err(pooled[0], oracle.fpn_forward(feats, rois, lv, scales, 7)) <= 1e-5


def test_errors_are_loud():
    x = torch.zeros(1, 4, 8, 8).cuda()
    r = torch.zeros(2, 5).cuda()
    with pytest.raises(_lib.RpoolError):
        ROIAlign2D(2, 2, 1.0, sampling_ratio=100).forward_gpu((x, r))
    with pytest.raises(TypeError):
        roi_align_2d(x.double(), r, 2, 2, 1.0)
    with pytest.raises(TypeError):
        roi_align_2d(x, r[:, :4], 2, 2, 1.0)
    with pytest.raises(TypeError):
        roi_align_2d(x.cpu(), r, 2, 2, 1.0)


@pytest.mark.gpu
def test_restated_cupy_kernels_agree_with_the_numpy_reference():
    """The same-box GPU baseline (baseline/refgpu_baseline.cu restates the reference's
    CuPy kernels, roi_align_2d.py:100-144 / :196-279) must compute the reference's
    op: forward within 1e-5 of the NumPy-path oracle; backward within 1e-4 on RoIs
    whose taps never coincide (the CuPy backward drops coincident-cell taps,
    :256-272, the NumPy backward does not)."""
    from baseline import refgpu
    rng = np.random.RandomState(5)
    x = rng.standard_normal((2, 8, 40, 56)).astype(np.float32)
    rois_yx = synth.make_rois(rng, 2, 24, 160, 224, size_range=(40.0, 150.0))
    # keep boxes away from the border so that x1 = min(x0 + 1, W - 1) never clamps
    rois_yx[:, 1:3] = np.maximum(rois_yx[:, 1:3], 8.0)
    rois_yx[:, 3] = np.minimum(rois_yx[:, 3], 150.0)
    rois_yx[:, 4] = np.minimum(rois_yx[:, 4], 214.0)
    rois_xy = oracle.roi_yx_to_xy(rois_yx)
    xd, rd = torch.from_numpy(x).cuda(), torch.from_numpy(rois_xy).cuda()
    top = refgpu.forward(xd, rd, 7, 7, 0.25)
    want = oracle.forward_chainer(x, rois_xy, 7, 7, 0.25)
    # 2e-5: the CuPy kernel forms the bin centre in double ((ph + 0.5) * bin + start, :127-128),
    # the NumPy path in float32 -- a 1-ulp coordinate difference
    assert oracle.rel_err(top.cpu().numpy(), want) <= 2e-5
    gy = rng.uniform(-1, 1, (rois_xy.shape[0], 8, 7, 7)).astype(np.float32)
    gx = refgpu.backward(torch.from_numpy(gy).cuda(), rd, x.shape, 0.25)
    want_g = oracle.backward_chainer(gy, rois_xy, x.shape, 0.25)
    assert oracle.rel_err(gx.cpu().numpy(), want_g) <= 1e-4
    # the per-RoI dispatch over a pyramid reproduces the batched calls
    feats = [xd, torch.from_numpy(rng.standard_normal((2, 8, 20, 28)).astype(np.float32)).cuda()]
    state = refgpu.FpnState(feats, [0.25, 0.125])
    levels = (np.arange(rois_xy.shape[0]) % 2).astype(np.int32)
    out = torch.empty((rois_xy.shape[0], 8, 7, 7), device="cuda")
    ops = refgpu.fpn_step(state, rd, levels, 7, out, torch.from_numpy(gy).cuda())
    torch.cuda.synchronize()
    assert ops == rois_xy.shape[0] * 4
    for l, sc in enumerate([0.25, 0.125]):
        m = np.nonzero(levels == l)[0]
        w = oracle.forward_chainer(feats[l].cpu().numpy(), rois_xy[m], 7, 7, sc)
        assert oracle.rel_err(out[torch.from_numpy(m).cuda()].cpu().numpy(), w) <= 2e-5
        wg = oracle.backward_chainer(gy[m], rois_xy[m], tuple(feats[l].shape), sc)
        assert oracle.rel_err(state.grads[l].cpu().numpy(), wg) <= 1e-4
