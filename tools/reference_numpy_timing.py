"""Times the reference's OWN, unmodified NumPy path (ROIAlign2D.forward_cpu / backward_cpu,
roi_align_2d.py:39-88,148-190, imported from /root/reference through
oracle/reference_loader.py) on bounded samples of BASELINE.json configs 0 and 1.

/root/reference exists only in the development container, so this cannot run on the
GPU box: the result is committed as profiles/r01_reference_numpy_cpu.json and quoted by
bench.py next to the CPU port it times live.  One op call per level (the
most favourable batching the reference API allows); 1 core, and all cores with one
process per shard of RoIs (dense gradients summed by the parent).
"""
import json
import multiprocessing as mp
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import synth
import oracle
from oracle import reference_loader

_STATE = {}


def _setup(cfg_id, n_sample):
    cfg = synth.CONFIGS[cfg_id]
    rng = np.random.RandomState(cfg_id)
    L = cfg["n_levels"]
    shapes = synth.pyramid_shapes(cfg["n_images"], cfg["channels"], cfg["height"], cfg["width"], L)
    rois = synth.make_rois(rng, cfg["n_images"], cfg["rois_per_image"], cfg["height"], cfg["width"],
                           aspect_range=cfg["aspect"])
    sel = np.sort(np.random.RandomState(99).choice(rois.shape[0], min(n_sample, rois.shape[0]), replace=False))
    rois = rois[sel]
    feats = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    levels = oracle.levels_for_pyramid(rois[:, 1:], L)
    scales = [1.0 / s for s in synth.STRIDES[:L]]
    P = cfg["out_sizes"][-1]
    gy = rng.uniform(-1, 1, (rois.shape[0], cfg["channels"], P, P)).astype(np.float32)
    return cfg, feats, rois, levels, scales, P, gy


def _run(feats, rois, levels, scales, P, gy):
    """forward + backward through the reference op, one call per level."""
    mod = reference_loader.load_reference_op()
    rois_xy = oracle.roi_yx_to_xy(rois)
    t_f = t_b = 0.0
    for l, x in enumerate(feats):
        sel = np.nonzero(levels == l)[0]
        if not sel.size:
            continue
        f = mod.ROIAlign2D(P, P, scales[l])
        r = np.ascontiguousarray(rois_xy[sel])
        t0 = time.perf_counter()
        f.forward_cpu((x, r))
        t1 = time.perf_counter()
        f._bottom_data_shape = x.shape
        f.backward_cpu((x, r), (np.ascontiguousarray(gy[sel]),))
        t2 = time.perf_counter()
        t_f += t1 - t0
        t_b += t2 - t1
    return t_f, t_b


def _worker(args):
    cfg_id, n_sample, k, n = args
    if cfg_id not in _STATE:
        _STATE[cfg_id] = _setup(cfg_id, n_sample)
    cfg, feats, rois, levels, scales, P, gy = _STATE[cfg_id]
    idx = np.arange(k, rois.shape[0], n)
    return _run(feats, rois[idx], levels[idx], scales, P, gy[idx])


def main():
    out = {"where": "development container (the reference tree cannot travel to the GPU box)",
           "cpu": platform.processor() or platform.machine(), "logical_cores": os.cpu_count(),
           "numpy": np.__version__, "python": platform.python_version(),
           "code": "unmodified chainer_maskrcnn/functions/roi_align/roi_align_2d.py NumPy bodies, "
                   "one call per level, 1 sample per bin (the reference path has no sampling_ratio)",
           "configs": {}}
    for cfg_id, n_sample in ((0, 512), (1, 192)):
        cfg, feats, rois, levels, scales, P, gy = _setup(cfg_id, n_sample)
        t_f, t_b = _run(feats, rois, levels, scales, P, gy)
        one = rois.shape[0] / (t_f + t_b)
        n = os.cpu_count() or 1
        with mp.Pool(n) as pool:
            pool.map(_worker, [(cfg_id, n_sample, k, n) for k in range(n)])          # warm-up: setup per worker
            t0 = time.perf_counter()
            pool.map(_worker, [(cfg_id, n_sample, k, n) for k in range(n)])
            wall = time.perf_counter() - t0
        out["configs"][cfg["name"]] = {
            "sample_rois": int(rois.shape[0]), "out": P,
            "one_core": {"fwd_s": t_f, "bwd_s": t_b, "rois_per_s": one},
            "all_cores": {"processes": n, "wall_s": wall, "rois_per_s": rois.shape[0] / wall,
                          "note": "RoIs dealt to one process per core; excludes summing the per-process "
                                  "dense gradients"},
        }
        print(cfg["name"], out["configs"][cfg["name"]], flush=True)
    path = os.path.join(ROOT, "profiles", "r01_reference_numpy_cpu.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
