"""Import the UNMODIFIED reference op from /root/reference.  TEST INFRASTRUCTURE ONLY.

``roi_align_2d.py`` imports ``chainer`` at module top (:9-12) but its NumPy
bodies (``forward_cpu`` :48-88, ``backward_cpu`` :149-190) only use ``numpy`` and
``six``.  chainer is not installed in this image, so a minimal stub module tree
is planted in ``sys.modules`` for the duration of the import.  The optional C++
extension is blocked (``caffe2_roi_align = None``) so ``forward_cpu`` cannot
silently switch to it (:41-46).

Works only where the reference tree exists (this container).  The GPU box has
no /root/reference: tests that need this module skip there and rely on the
golden vectors in tests/golden/ generated from it (tests/golden/make_golden.py).
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RPOOL_REFERENCE_ROOT", "/root/reference")
_OP_FILE = os.path.join(REFERENCE_ROOT, "chainer_maskrcnn", "functions", "roi_align",
                        "roi_align_2d.py")
_RPN_FILE = os.path.join(REFERENCE_ROOT, "chainer_maskrcnn", "model", "rpn",
                         "multilevel_region_proposal_network.py")


def set_root(root):
    """Points the loader at another copy of the reference files (baseline/_ref on the GPU
    box, installed byte for byte by `make -C baseline ref`)."""
    global REFERENCE_ROOT, _OP_FILE, _RPN_FILE, _cached
    REFERENCE_ROOT = root
    _OP_FILE = os.path.join(root, "chainer_maskrcnn", "functions", "roi_align", "roi_align_2d.py")
    _RPN_FILE = os.path.join(root, "chainer_maskrcnn", "model", "rpn",
                             "multilevel_region_proposal_network.py")
    _cached = None


def available():
    return os.path.exists(_OP_FILE)


def _stub_chainer():
    chainer = types.ModuleType("chainer")
    cuda = types.ModuleType("chainer.cuda")
    function = types.ModuleType("chainer.function")
    function_node = types.ModuleType("chainer.function_node")
    utils = types.ModuleType("chainer.utils")
    type_check = types.ModuleType("chainer.utils.type_check")

    class Function(object):
        pass

    function.Function = Function
    chainer.cuda, chainer.function, chainer.function_node, chainer.utils = (
        cuda, function, function_node, utils)
    utils.type_check = type_check
    return {
        "chainer": chainer, "chainer.cuda": cuda, "chainer.function": function,
        "chainer.function_node": function_node, "chainer.utils": utils,
        "chainer.utils.type_check": type_check,
    }


_cached = None


def load_reference_op():
    """Returns the reference module object (attributes ROIAlign2D, roi_align_2d)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    stubs = _stub_chainer()
    stubs["caffe2_roi_align"] = None  # block the optional C++ port
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("_reference_roi_align_2d", _OP_FILE)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if k == "caffe2_roi_align":
                continue  # stays blocked: forward_cpu re-imports it on every call
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cached = mod
    return mod


def reference_forward(x, rois, outh, outw, spatial_scale):
    """ROIAlign2D(outh,outw,scale).forward_cpu((x, rois))[0], unmodified reference."""
    mod = load_reference_op()
    return mod.ROIAlign2D(outh, outw, spatial_scale).forward_cpu((x, rois))[0]


def reference_backward(gy, x, rois, outh, outw, spatial_scale):
    """backward_cpu after a forward_cpu (which records the input shape, :40)."""
    mod = load_reference_op()
    f = mod.ROIAlign2D(outh, outw, spatial_scale)
    f._bottom_data_shape = x.shape
    return f.backward_cpu((x, rois), (gy,))[0]


def reference_level_mapper():
    """map_rois_to_fpn_levels (multilevel_region_proposal_network.py:16-31) executed
    from the reference file itself: the function's source lines are compiled in
    isolation (the module's other imports need chainercv) with ``chainer``'s
    get_array_module answering numpy."""
    import numpy as np
    with open(_RPN_FILE) as f:
        lines = f.read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("def map_rois_to_fpn_levels"))
    end = next(i for i in range(start + 1, len(lines))
               if lines[i] and not lines[i].startswith((" ", "\t")))
    src = "\n".join(lines[start:end])
    chainer = types.SimpleNamespace(backends=types.SimpleNamespace(
        cuda=types.SimpleNamespace(get_array_module=lambda *a: np)))
    ns = {"chainer": chainer, "np": np}
    exec(compile(src, _RPN_FILE, "exec"), ns)
    return ns["map_rois_to_fpn_levels"]
