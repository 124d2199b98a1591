"""CPU: the C-ABI library loads, exports every symbol include/rpool_b200.h
declares, agrees with the ctypes struct layout, and rejects bad arguments
without touching a GPU (no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pytest

import chainer_maskrcnn_b200 as pkg
from chainer_maskrcnn_b200 import _lib, _engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rpool_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"RPOOL_API[^;(]*?\b(rpool_[a-z0-9_]+)\s*\(", text)))


def test_library_is_in_tree_and_loads():
    L = _lib.lib()
    assert os.path.dirname(_lib._build.LIB_PATH) == os.path.join(ROOT, "chainer-maskrcnn_b200")
    assert L.rpool_version() == 200 == _lib.VERSION


def test_every_declared_symbol_is_exported():
    L = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(_lib.EXPORTS) == declared


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "rpool_b200.h"\nint main(void){return (int)sizeof(rpool_problem)==0;}\n')
    import subprocess
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.dirname(HEADER),
                           "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_struct_layout_matches():
    assert _lib.lib().rpool_problem_size() == ctypes.sizeof(_lib.Problem)


def test_thresholds_libc_equal_numpy():
    for kmin, kmax in [(0, 4), (0, 3), (1, 4), (2, 2)]:
        a = _lib.level_thresholds_libc(k_min=kmin, k_max=kmax)
        b = _engine.level_thresholds(kmin, kmax)
        assert np.array_equal(np.array(a, np.float32), np.array(b, np.float32))


def test_build_id_matches_the_sources():
    """A stale binary is never loaded: the library carries the hash of the sources it was
    compiled from, and the binding compares it with the tree's."""
    assert _lib.build_id() == _lib._build.source_hash() != "unknown"
    assert not _lib._build.is_stale()


def test_binding_needs_nothing_but_the_standard_library():
    """`import chainer_maskrcnn_b200._lib` must not pull torch or numpy: a CuPy/Chainer host
    binds the C ABI through it."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import chainer_maskrcnn_b200._lib as L; "
            "assert L.lib().rpool_version() == 200; "
            "bad = [m for m in ('torch', 'numpy', 'cupy') if m in sys.modules]; assert not bad, bad" % ROOT)
    subprocess.check_call([sys.executable, "-c", code])


def test_options_validation():
    L = _lib.lib()
    raw = ctypes.create_string_buffer(65536 + 16)
    ws = ctypes.c_void_p((ctypes.addressof(raw) + 15) & ~15)
    for bad in (dict(cta_threads=100), dict(cta_threads=512), dict(schedule=4), dict(force_path=3),
                dict(prefetch_rows=17), dict(prefetch_rois=-1), dict(fuse_heads_backward=2),
                dict(backward_variant=3)):
        p = _problem()
        p.opt = _lib.make_options(**bad)
        for fn in (L.rpool_plan, L.rpool_forward, L.rpool_backward):
            assert fn(ctypes.byref(p), ws, 65536, None) == 1, bad
            assert b"opt." in L.rpool_last_error()
    with pytest.raises(ValueError):
        _lib.make_options(nonsense=1)


def _problem(**kw):
    p = _lib.Problem()
    p.n_levels = 1
    p.channels = 8
    p.level[0].data = 0x1000
    p.level[0].n_images = 1
    p.level[0].height = 4
    p.level[0].width = 4
    p.level[0].spatial_scale = 1.0
    p.rois = 0x2000
    p.n_rois = 4
    p.n_heads = 1
    p.out_h[0] = p.out_w[0] = 2
    p.pooled[0] = 0x3000
    p.sampling_ratio = 1
    for k, v in kw.items():
        setattr(p, k, v)
    return p


@pytest.mark.parametrize("kw,code", [
    (dict(n_levels=0), 1), (dict(n_levels=9), 1), (dict(channels=0), 1),
    (dict(feat_layout=7), 1), (dict(roi_format=3), 1), (dict(n_rois=-1), 1),
    (dict(rois=None), 1), (dict(n_heads=3), 1), (dict(sampling_ratio=2), 1),
    (dict(coord_mode=5), 1), (dict(sampling_ratio=100, coord_mode=1), 2),
])
def test_validation_errors_without_gpu(kw, code):
    L = _lib.lib()
    p = _problem(**kw)
    raw = ctypes.create_string_buffer(65536 + 16)
    ws = ctypes.c_void_p((ctypes.addressof(raw) + 15) & ~15)
    for fn in (L.rpool_plan, L.rpool_forward, L.rpool_backward):
        rc = fn(ctypes.byref(p), ws, 65536, None)
        assert rc == code, (fn.__name__, kw, L.rpool_last_error())
        assert len(L.rpool_last_error()) > 0


def test_workspace_errors_without_gpu():
    L = _lib.lib()
    p = _problem()
    assert L.rpool_forward(ctypes.byref(p), None, 0, None) == 3
    ws = ctypes.create_string_buffer(16)
    assert L.rpool_forward(ctypes.byref(p), ws, 16, None) == 3
    assert L.rpool_workspace_bytes(1000) >= 3 * 4 * 1000
    assert L.rpool_plan(None, ws, 16, None) == 1
    assert L.rpool_zero_fill(None, None) == 1
    q = _problem()
    q.level[0].data = None
    assert L.rpool_zero_fill(ctypes.byref(q), None) == 1
    # the deterministic variant needs its scratch, channels-last tensors and the default schedule
    p = _problem(deterministic=1)
    n = L.rpool_workspace_bytes_ex(p.n_rois, p.n_heads, p.coord_mode)
    assert 0 < n <= L.rpool_workspace_bytes(p.n_rois)
    raw = ctypes.create_string_buffer(n + 16)
    ws = ctypes.c_void_p((ctypes.addressof(raw) + 15) & ~15)
    assert L.rpool_backward(ctypes.byref(p), ws, n - 1, None) == 3
    assert b"needed" in L.rpool_last_error()
    assert L.rpool_backward(ctypes.byref(p), ctypes.c_void_p(ws.value + 4), n, None) == 3
    assert b"aligned" in L.rpool_last_error()
    assert L.rpool_backward(ctypes.byref(p), ws, n, None) == 3
    assert b"det_workspace" in L.rpool_last_error()
    p = _problem(deterministic=1, det_workspace=0x4000, det_workspace_bytes=1 << 20, feat_layout=1)
    assert L.rpool_backward(ctypes.byref(p), ws, n, None) == 2
    assert b"channels-last" in L.rpool_last_error()
    p = _problem(deterministic=1, det_workspace=0x4000, det_workspace_bytes=1 << 20)
    p.opt = _lib.make_options(schedule=_lib.SCHED_INPUT)
    assert L.rpool_backward(ctypes.byref(p), ws, n, None) == 2
    assert b"schedule" in L.rpool_last_error()


def test_launch_without_gpu_fails_loudly():
    """No CPU fallback: with no device the launch itself must report an error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _lib.lib()
    thr = (ctypes.c_float * 4)(1, 2, 3, 4)
    rc = L.rpool_assign_levels(0x1000, 4, 4, 1, thr, 4, 0, 4, 0x2000, None, None)
    assert rc == 4 and b":" in L.rpool_last_error()
    with pytest.raises(TypeError):
        pkg.roi_align_2d(torch.zeros(1, 4, 8, 8), torch.zeros(1, 5), 2, 2, 1.0)
    with pytest.raises(RuntimeError):
        pkg.roi_align_2d(np.zeros((1, 4, 8, 8), np.float32), np.zeros((1, 5), np.float32), 2, 2, 1.0)
