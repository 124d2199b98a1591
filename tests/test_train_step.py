"""cfg 5: the torch restatement of the reference training step
(model/fpn_maskrcnn_train_chain.py) around the pooling path.

CPU: the step runs, every loss is finite, parameters move, the per-RoI dispatch of the
reference (fpn_roi_mask_head.py:57-63) and the per-level batched dispatch give the same
loss.  GPU: swapping this package's fused kernels in gives the same loss and gradients
as torchvision's roi_align (caffe2 semantics on both sides)."""
import pytest
import torch

from chainer_maskrcnn_b200.model.fpn_maskrcnn_train_chain import (
    AnchorTargetCreator, FPNMaskRCNNTrainChain, MaskRCNN, ProposalTargetCreator, bbox_iou,
    synthetic_batch)
from chainer_maskrcnn_b200.model.extractor.feature_pyramid_network import FeaturePyramidNetwork
from chainer_maskrcnn_b200.model.rpn.multilevel_region_proposal_network import (
    bbox2loc, generate_anchor_base, levels_by_formula, loc2bbox)

THIN = dict(width=8, blocks=(1, 1, 1, 1), channels=16, fc_dim=32,
            proposal_creator_params=dict(n_train_pre_nms=500, n_train_post_nms=100))


def _chain(pooling, device, level_fn, sampling_ratio=2, seed=0):
    torch.manual_seed(seed)
    model = MaskRCNN(5, pooling=pooling, sampling_ratio=sampling_ratio, level_fn=level_fn, **THIN)
    return FPNMaskRCNNTrainChain(model, level_fn=level_fn, seed=seed).to(device)


def test_pyramid_shapes_follow_the_reference_extractor():
    # feature_pyramid_network.py:48-53: conv1 s2 p3 -> cover-all max-pool -> stride-2 stages
    fpn = FeaturePyramidNetwork(width=4, blocks=(1, 1, 1, 1), out_channels=8).eval()
    with torch.no_grad():
        ps = fpn(torch.zeros(1, 3, 100, 167))
    assert [tuple(p.shape[2:]) for p in ps] == [(25, 42), (13, 21), (7, 11), (4, 6), (2, 3)]
    assert fpn.spatial_scales == [1 / 4, 1 / 8, 1 / 16, 1 / 32, 1 / 64]


def test_box_coding_round_trip_and_anchor_base():
    g = torch.Generator().manual_seed(1)
    src = torch.rand(50, 4, generator=g) * 100
    src[:, 2:] += src[:, :2] + 5
    dst = torch.rand(50, 4, generator=g) * 100
    dst[:, 2:] += dst[:, :2] + 5
    assert torch.allclose(loc2bbox(src, bbox2loc(src, dst)), dst, atol=1e-3)
    base = generate_anchor_base(anchor_scales=[2.0])
    assert base.shape == (3, 4)
    area = (base[:, 2] - base[:, 0]) * (base[:, 3] - base[:, 1])
    assert torch.allclose(area, torch.full((3,), 32.0 * 32.0), rtol=1e-5)   # anchor_sizes[0] = 32


def test_target_creators_counts():
    g = torch.Generator().manual_seed(0)
    bbox = torch.tensor([[10., 10., 60., 80.], [30., 90., 120., 150.]])
    roi = torch.rand(300, 4, generator=g) * 100
    roi[:, 2:] = roi[:, :2] + 10 + torch.rand(300, 2, generator=g) * 60
    lv = levels_by_formula(roi)
    mask = torch.zeros(2, 160, 160, dtype=torch.bool)
    mask[0, 10:60, 10:80] = True
    mask[1, 30:120, 90:150] = True
    ptc = ProposalTargetCreator(n_sample=64, level_fn=levels_by_formula)
    s_roi, s_lv, loc, label, m = ptc(roi, bbox, torch.tensor([1, 3]), mask, lv, mask_size=28)
    assert s_roi.shape[0] == s_lv.shape[0] == loc.shape[0] == label.shape[0] <= 64
    n_pos = int((label > 0).sum())
    assert 2 <= n_pos <= 16 and m.shape == (n_pos, 28, 28)            # GT boxes are appended (:46)
    assert bool((label[:n_pos] > 0).all()) and bool((label[n_pos:] == 0).all())
    # the appended GT boxes keep their own levels (:49-50)
    assert torch.equal(levels_by_formula(s_roi), s_lv)
    atc = AnchorTargetCreator(n_sample=32)
    anchor = torch.cat([roi, bbox])
    a_loc, a_label = atc(bbox, anchor, (160, 160))
    assert a_loc.shape == (302, 4) and int((a_label >= 0).sum()) <= 32
    assert float(bbox_iou(bbox, bbox).diagonal().min()) == pytest.approx(1.0)


def test_train_step_cpu_runs_and_dispatches_agree():
    dev = torch.device("cpu")
    imgs, b, l, k = synthetic_batch(2, 160, 224, 5, 4, seed=3)
    losses = {}
    for pooling in ("torchvision", "per_roi"):
        chain = _chain(pooling, dev, levels_by_formula)
        opt = torch.optim.SGD(chain.parameters(), lr=1e-2, momentum=0.9)
        w0 = chain.faster_rcnn.head.fc1.weight.detach().clone()
        loss = chain(imgs, b, l, k)
        opt.zero_grad()
        loss.backward()
        opt.step()
        for name in ("rpn_loc_loss", "rpn_cls_loss", "roi_loc_loss", "roi_cls_loss", "mask_loss"):
            assert torch.isfinite(chain.last[name]), name
        assert not torch.equal(w0, chain.faster_rcnn.head.fc1.weight)
        assert chain.faster_rcnn.extractor.conv1.weight.grad.abs().sum() > 0   # pooling passes gradients on
        losses[pooling] = float(loss.detach())
    assert losses["torchvision"] == pytest.approx(losses["per_roi"], rel=1e-5)


def test_b200_pooling_refuses_cpu_tensors():
    chain = _chain("b200", torch.device("cpu"), levels_by_formula)
    imgs, b, l, k = synthetic_batch(1, 96, 128, 5, 3, seed=1)
    with pytest.raises((TypeError, RuntimeError)):
        chain(imgs, b, l, k)


@pytest.mark.gpu
def test_train_step_b200_pooling_matches_torchvision():
    dev = torch.device("cuda", 0)
    imgs, b, l, k = synthetic_batch(2, 256, 320, 5, 4, seed=3, device=dev)
    out = {}
    for pooling in ("b200", "torchvision"):
        # levels from the device mapper on both sides: identical sampled RoIs
        chain = _chain(pooling, dev, None).to(memory_format=torch.channels_last)
        loss = chain(imgs.contiguous(memory_format=torch.channels_last), b, l, k)
        loss.backward()
        m = chain.faster_rcnn
        out[pooling] = (float(loss.detach()), m.extractor.conv_p2.weight.grad.clone(),
                        m.extractor.conv1.weight.grad.clone(), chain.last["n_sample"])
    assert out["b200"][3] == out["torchvision"][3]
    assert out["b200"][0] == pytest.approx(out["torchvision"][0], rel=1e-5)
    # the lateral conv right under the pooled level sees the pooling gradient almost directly; conv1's
    # weight gradient went through the whole backbone backwards (atomics in both pooling back ends, cuDNN
    # reductions): it agrees to ~8e-4 of its largest entry and moves by 1e-4 from run to run
    for i, tol in ((1, 1e-4), (2, 5e-3)):
        a, r = out["b200"][i], out["torchvision"][i]
        assert float((a - r).abs().max() / r.abs().max()) <= tol
