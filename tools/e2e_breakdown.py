"""Where the end-to-end (host arrays in, host arrays out) time goes: each phase
of fpn_roi_align_host timed alone on cfg1."""
import json, sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, synth
import chainer_maskrcnn_b200 as pkg
from chainer_maskrcnn_b200 import _host, _engine, _lib

cfg = synth.CONFIGS[1]
rng = np.random.RandomState(1)
shapes = synth.pyramid_shapes(cfg["n_images"], 256, cfg["height"], cfg["width"], 4)
rois = synth.make_rois(rng, cfg["n_images"], cfg["rois_per_image"], cfg["height"], cfg["width"])
scales = [1 / s for s in synth.STRIDES[:4]]
pin = lambda shape: torch.randn(shape).pin_memory().numpy()
feats = [pin(s) for s in shapes]
gy = pin((rois.shape[0], 256, 14, 14))
rois_p = torch.from_numpy(rois).pin_memory().numpy()
def T(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
res = {}
res["h2d_feats_ms"] = T(lambda: [_host.h2d(f) for f in feats])
res["h2d_gy_ms"] = T(lambda: _host.h2d(gy))
fd = [_host.h2d(f) for f in feats]; gd = _host.h2d(gy); rd = _host.h2d(rois_p)
res["fwd_incl_nchw2nhwc_ms"] = T(lambda: _engine.forward(fd, rd, None, scales, [14], 2))
outs, plan = _engine.forward(fd, rd, None, scales, [14], 2)
res["bwd_incl_gy_transpose_ms"] = T(lambda: _engine.backward(plan, [gd]))
res["d2h_pooled_ms"] = T(lambda: _host.d2h(outs[0]))
def full():
    p, g = pkg.fpn_roi_align_host(feats, rois_p, None, scales, [14], 2, gys=[gy]); del p, g
res["full_ms"] = T(full)
def full1():
    p, g = pkg.fpn_roi_align_host(feats, rois_p, None, scales, [14], 2, gys=[gy], max_groups=1); del p, g
res["full_single_group_ms"] = T(full1)
res["full_ms_20"] = T(full, 20)
res["full_single_group_ms_20"] = T(full1, 20)
res["full_ms_20_again"] = T(full, 20)
print(json.dumps(res))
