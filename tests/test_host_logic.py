"""CPU: host-side mirror of the reference interface (names, signatures, type
checks) and the multi-GPU partitioning, including a world_size-2 gloo run."""
import inspect
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import chainer_maskrcnn_b200 as pkg
from chainer_maskrcnn_b200 import _engine, _sharding
from chainer_maskrcnn_b200.functions.roi_align.roi_align_2d import ROIAlign2D, InvalidType, roi_align_2d
from chainer_maskrcnn_b200.functions.roi_align_2d_yx import _roi_align_2d_yx
from chainer_maskrcnn_b200.model.rpn.multilevel_region_proposal_network import map_rois_to_fpn_levels
from chainer_maskrcnn_b200.model.extractor import feature_pyramid_network as fpn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_names_and_signatures():
    # roi_align_2d.py:18, :284 ; roi_align_2d_yx.py:4 ; multilevel_region_proposal_network.py:16
    assert list(inspect.signature(ROIAlign2D.__init__).parameters)[:4] == \
        ["self", "outh", "outw", "spatial_scale"]
    assert list(inspect.signature(roi_align_2d).parameters)[:5] == \
        ["x", "rois", "outh", "outw", "spatial_scale"]
    assert list(inspect.signature(_roi_align_2d_yx).parameters)[:5] == \
        ["x", "indices_and_rois", "outh", "outw", "spatial_scale"]
    sig = inspect.signature(map_rois_to_fpn_levels)
    assert list(sig.parameters) == ["rois", "k_min", "k_max"]
    assert sig.parameters["k_min"].default == 0 and sig.parameters["k_max"].default == 4
    for m in ("check_type_forward", "forward_cpu", "forward_cpu2", "forward_gpu", "backward_cpu",
              "backward_gpu"):
        assert callable(getattr(ROIAlign2D, m))
    assert list(inspect.signature(pkg.FPNRoIPooling.__call__).parameters)[:5] == \
        ["self", "x", "indices_and_rois", "levels", "spatial_scales"]
    assert list(inspect.signature(pkg.FPNRoIPooling.predict_mask).parameters) == \
        ["self", "levels", "indices_and_rois", "spatial_scales"]


def test_chainer_adapter_is_import_guarded():
    """chainer / cupy are absent here: the adapter says so instead of half-importing."""
    import importlib
    import sys
    if "chainer" in sys.modules or importlib.util.find_spec("chainer") is not None:
        pytest.skip("chainer is installed")
    with pytest.raises(ImportError, match="needs chainer and cupy"):
        importlib.import_module("chainer_maskrcnn_b200.chainer_adapter")


def test_caffe2_module_shim_has_the_reference_signature():
    # caffe2_roi_align.cpp:231: forward(bottom_data, bottom_rois, out_h, out_w, spatial_scale)
    import importlib.util
    path = os.path.join(ROOT, "chainer-maskrcnn_b200", "dropin", "caffe2_roi_align.py")
    spec = importlib.util.spec_from_file_location("_shim_caffe2_roi_align", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert list(inspect.signature(mod.forward).parameters) == \
        ["bottom_data", "bottom_rois", "out_h", "out_w", "spatial_scale"]
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):          # no device: loud, never a CPU result
            mod.forward(np.zeros((1, 4, 8, 8), np.float32), np.zeros((1, 5), np.float32), 2, 2, 1.0)


def test_pyramid_constants():
    assert fpn.feat_strides == [4, 8, 16, 32, 64]          # feature_pyramid_network.py:9
    assert fpn.spatial_scales == [1 / 4, 1 / 8, 1 / 16, 1 / 32, 1 / 64]
    assert fpn.pyramid_shapes(2, 256, 800, 1333, 4) == \
        [(2, 256, 200, 334), (2, 256, 100, 167), (2, 256, 50, 84), (2, 256, 25, 42)]


def test_check_type_forward():
    f = ROIAlign2D(7, 7, 0.25)
    x = np.zeros((1, 4, 8, 8), np.float32)
    r = np.zeros((3, 5), np.float32)
    f.check_type_forward((x, r))
    f.check_type_forward((torch.zeros(1, 4, 8, 8), torch.zeros(3, 5)))
    for bad in [(x,), (x.astype(np.float64), r), (x[0], r), (x, r.astype(np.int32)),
                (x, r[:, :4]), (x, r[0])]:
        with pytest.raises(InvalidType):
            f.check_type_forward(bad)


def test_level_thresholds_are_cached_monotone_and_reference_exact():
    t = _engine.level_thresholds()
    assert t is _engine.level_thresholds()
    assert list(np.array(t, np.float32).view(np.uint32)) == \
        [0x4443ff31, 0x4543ff98, 0x4643ffc9, 0x4743ffe4]
    assert all(b > a for a, b in zip(t, t[1:]))


def test_out_size_normalisation():
    assert _engine._norm_sizes([7]) == [(7, 7)]
    assert _engine._norm_sizes([(5, 7), 14]) == [(5, 7), (14, 14)]
    with pytest.raises(ValueError):
        _engine._norm_sizes([7, 14, 28])


def test_sharding_by_image():
    rois = np.zeros((10, 5), np.float32)
    rois[:, 0] = [0, 0, 1, 1, 2, 2, 3, 3, 4, 4]
    rois[:, 1] = np.arange(10)
    assert _sharding.images_of_rank(5, 2, 0) == [0, 2, 4]
    assert _sharding.images_of_rank(5, 2, 1) == [1, 3]
    seen = []
    for rank in range(2):
        local, rows = _sharding.shard_rois(rois, 5, 2, rank)
        assert np.array_equal(local[:, 1], rois[rows, 1])
        assert set(local[:, 0]) == set(range(len(_sharding.images_of_rank(5, 2, rank))))
        seen += list(rows)
    assert sorted(seen) == list(range(10))
    assert _sharding.max_over_ranks(3.5) == 3.5


def test_host_call_groups_images_only_when_rois_are_image_major():
    from chainer_maskrcnn_b200.functions.fpn_roi_align import _image_groups
    r = np.zeros((10, 5), np.float32)
    r[:, 0] = [0, 0, 0, 1, 1, 2, 2, 2, 2, 3]
    assert _image_groups(r, 4) == [(0, 1, 0, 3), (1, 2, 3, 5), (2, 3, 5, 9), (3, 4, 9, 10)]
    # more images than groups: contiguous image ranges; images without RoIs get empty row ranges
    g = _image_groups(r, 8)
    assert [x[:2] for x in g] == [(0, 2), (2, 4), (4, 6), (6, 8)]
    assert [x[2:] for x in g] == [(0, 5), (5, 10), (10, 10), (10, 10)]
    assert _image_groups(r, 4, max_groups=1) == [(0, 4, 0, 10)]
    assert _image_groups(r, 1) == [(0, 1, 0, 10)]
    assert _image_groups(r[:0], 4) == [(0, 4, 0, 0)]
    for bad in ([0, 1, 0, 1, 1, 2, 2, 2, 2, 3], [0, 0, 0, 1, 1, 2, 2, 2, 2, 4], [0, 0, 0, 1, 1, 2, 2, 2, 2, 2.5]):
        r2 = r.copy()
        r2[:, 0] = bad
        assert _image_groups(r2, 4) == [(0, 4, 0, 10)]        # interleaved / out of range / fractional


def test_channel_padding_helper_and_plan_fields():
    # channel counts that are not a multiple of 4 are zero-extended by the shim (the light
    # head's 490-channel map): layout, values and padding of the helper, on CPU tensors
    t = torch.randn(2, 6, 5, 7)
    o = _engine._pad_channels(t, 8)
    assert tuple(o.shape) == (2, 8, 5, 7) and o.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(o[:, :6], t) and float(o[:, 6:].abs().max()) == 0.0
    assert tuple(_engine._pad_channels(torch.randn(0, 6, 5, 7), 8).shape) == (0, 8, 5, 7)
    assert "cpad" in _engine.Plan.__slots__ and "problem" in _engine.Plan.__slots__
    with pytest.raises(TypeError):
        _engine.make_plan([(1, 6, 8, 8)], torch.zeros(3, 5), None, [0.25], [7])      # CPU RoIs: no fallback


def test_bench_sharded_workload_covers_every_roi_once():
    # bench.py --shard: ONE instance of the config, images dealt round-robin to the ranks
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth
    cfg = synth.CONFIGS[3]
    rng = np.random.RandomState(3)
    rois = synth.make_rois(rng, cfg["n_images"], cfg["rois_per_image"], cfg["height"], cfg["width"],
                           aspect_range=cfg["aspect"])
    for world in (2, 4, 8):
        seen, n_img = [], 0
        for rank in range(world):
            local, rows = _sharding.shard_rois(rois, cfg["n_images"], world, rank)
            mine = _sharding.images_of_rank(cfg["n_images"], world, rank)
            n_img += len(mine)
            assert local[:, 0].max() == len(mine) - 1 and local[:, 0].min() == 0
            assert np.array_equal(local[:, 1:], rois[rows, 1:])
            seen.append(rows)
        assert n_img == cfg["n_images"]
        assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(rois.shape[0]))


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from chainer_maskrcnn_b200 import _sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rois = np.zeros((12, 5), np.float32); rois[:, 0] = np.repeat(np.arange(6), 2)
local, rows = _sharding.shard_rois(rois, 6, world, rank)
n = _sharding.sum_over_ranks(len(rows))
t = _sharding.max_over_ranks(1.0 + rank)
assert n == 12, n
assert t == float(world), t
assert sorted(set(local[:, 0])) == [0, 1, 2]
dist.barrier()
if rank == 0:
    print("GLOO_OK", n, t)
dist.destroy_process_group()
"""


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
        capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO_OK 12.0 2.0" in out.stdout
