"""The reference's OWN NumPy path, unmodified, timed on this machine's host cores.
BENCH INFRASTRUCTURE ONLY -- never imported by the product.

`ROIAlign2D.forward_cpu` / `backward_cpu` (roi_align_2d.py:39-88,148-190) are imported
from the installed copy in baseline/_ref/ (`make -C baseline ref`; byte-for-byte
files, sha256 in baseline/_ref/MANIFEST) -- or from /root/reference where that
exists -- through oracle/reference_loader.py (stub `chainer`, C++ extension blocked).
The path is pure-Python loops over RoI x bin, so it is single-threaded: the
all-core figure runs one process per core over RoI shards (`fork`, the pyramid is
shared copy-on-write), which is the most favourable way to use the box for it.
One op call per level (the batching the reference API allows); the reference path
samples once per bin (it has no sampling_ratio).
"""
import multiprocessing as mp
import os
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
_JOB = {}


def build(force=False):
    """Installs the reference files into baseline/_ref when the reference tree is here."""
    import subprocess
    if os.path.isdir("/root/reference/chainer_maskrcnn") and (force or not available_installed()):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def available_installed():
    return os.path.exists(os.path.join(REF_DIR, "chainer_maskrcnn", "functions", "roi_align",
                                       "roi_align_2d.py"))


def _loader():
    """oracle.reference_loader pointed at the installed copy (or the reference tree)."""
    from oracle import reference_loader as rl
    if not rl.available() and available_installed():
        rl.set_root(REF_DIR)
    return rl


def available():
    return _loader().available()


def source():
    rl = _loader()
    return "baseline/_ref (installed by `make -C baseline ref`)" if rl.REFERENCE_ROOT == REF_DIR \
        else rl.REFERENCE_ROOT


def run_levels(feats, rois_xy, levels, scales, P, gy):
    """forward + backward through the reference op, one call per level.
    Returns (fwd_seconds, bwd_seconds)."""
    mod = _loader().load_reference_op()
    t_f = t_b = 0.0
    for l, x in enumerate(feats):
        sel = np.nonzero(levels == l)[0]
        if not sel.size:
            continue
        f = mod.ROIAlign2D(P, P, scales[l])
        r = np.ascontiguousarray(rois_xy[sel])
        g = np.ascontiguousarray(gy[sel])
        t0 = time.perf_counter()
        f.forward_cpu((x, r))
        t1 = time.perf_counter()
        f._bottom_data_shape = x.shape
        f.backward_cpu((x, r), (g,))
        t2 = time.perf_counter()
        t_f += t1 - t0
        t_b += t2 - t1
    return t_f, t_b


def _worker(k):
    j = _JOB
    idx = np.arange(k, j["rois_xy"].shape[0], j["n"])
    return run_levels(j["feats"], j["rois_xy"][idx], j["levels"][idx], j["scales"], j["P"], j["gy"][idx])


def time_path(feats, rois_yx, levels, scales, P, gy, rois_1core=48, rois_per_process=96, cores=None):
    """Times the path on bounded samples: `rois_1core` RoIs on one core, and
    cores x `rois_per_process` RoIs with one process per core.  Returns a dict."""
    cores = cores or len(os.sched_getaffinity(0))
    rois_xy = np.ascontiguousarray(rois_yx[:, [0, 2, 1, 4, 3]])
    R = rois_xy.shape[0]
    pick = np.random.RandomState(99).permutation(R)
    s1 = np.sort(pick[:min(rois_1core, R)])
    t_f, t_b = run_levels(feats, rois_xy[s1], levels[s1], scales, P, gy[s1])
    out = {"source": source(), "where": "this box (timed live by bench.py)",
           "one_core": {"rois": int(s1.size), "fwd_s": t_f, "bwd_s": t_b,
                        "rois_per_s": s1.size / (t_f + t_b)}}
    sn = np.sort(pick[:min(rois_per_process * cores, R)])
    _JOB.update(feats=feats, rois_xy=rois_xy[sn], levels=levels[sn], scales=scales, P=P, gy=gy[sn], n=cores)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_worker, range(cores), chunksize=1)              # import + first touch per worker
        t0 = time.perf_counter()
        pool.map(_worker, range(cores), chunksize=1)
        wall = time.perf_counter() - t0
    _JOB.clear()
    out["all_cores"] = {"processes": int(cores), "rois": int(sn.size), "wall_s": wall,
                        "rois_per_s": sn.size / wall,
                        "note": "one process per core over RoI shards; summing the per-process dense "
                                "gradients is not included"}
    return out
