"""Host-side cost of one fused call on a small workload (cfg 0): wall time per
_engine.forward / backward with the GPU kept idle-free, and a cProfile of the loop."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import synth
from chainer_maskrcnn_b200 import _engine, _lib

cfg = synth.CONFIGS[0]
rng = np.random.RandomState(0)
L = cfg["n_levels"]
shapes = synth.pyramid_shapes(cfg["n_images"], cfg["channels"], cfg["height"], cfg["width"], L)
rois = torch.from_numpy(synth.make_rois(rng, cfg["n_images"], cfg["rois_per_image"], cfg["height"],
                                        cfg["width"], aspect_range=cfg["aspect"])).cuda()
scales = [1.0 / s for s in synth.STRIDES[:L]]
feats = [torch.randn(s, device="cuda").contiguous(memory_format=torch.channels_last) for s in shapes]
sizes = cfg["out_sizes"]
gys = [torch.rand((rois.shape[0], cfg["channels"], P, P), device="cuda")
       .contiguous(memory_format=torch.channels_last) for P in sizes]
grads = [torch.empty(s, device="cuda").contiguous(memory_format=torch.channels_last) for s in shapes]


def loop(n):
    for _ in range(n):
        outs, plan = _engine.forward(feats, rois, None, scales, sizes, sampling_ratio=2)
        _engine.backward(plan, gys, out=grads)


loop(200)
torch.cuda.synchronize()
n = 2000
t0 = time.perf_counter()
loop(n)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host us per fwd+bwd call pair: %.1f   (after sync: %.1f)" % ((t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
pr = cProfile.Profile()
pr.enable()
loop(n)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(18)
