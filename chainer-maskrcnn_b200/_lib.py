"""ctypes binding of librpool_b200.so (include/rpool_b200.h).

Array-library agnostic: every call takes raw device addresses (ints) and a
stream handle.  There is no CPU fallback -- if the shared object is missing
(and cannot be built) or no CUDA device is present, calls raise.
"""
import ctypes
import os

from . import _build

MAX_LEVELS = 8
MAX_HEADS = 2
VERSION = 200

NHWC, NCHW = 0, 1
ROI_XY, ROI_YX = 0, 1
COORD_CHAINER, COORD_CAFFE2 = 0, 1
PATH_AUTO, PATH_GENERIC, PATH_TABLE = 0, 1, 2

SCHED_DEFAULT, SCHED_INPUT, SCHED_LEVEL_DESC, SCHED_COARSE_FIRST = 0, 1, 2, 3
FLAG_BAD_BATCH, FLAG_LEVEL_CLIPPED, FLAG_DET_GENERIC, FLAG_DET_SCRATCH = 1, 2, 4, 8

UNSUPPORTED = 2
_STATUS = {1: "invalid argument", 2: "unsupported", 3: "workspace", 4: "CUDA error"}


class RpoolError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("librpool_b200: %s (%s)" % (msg, _STATUS.get(code, code)))
        self.code = code


class Level(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p),
                ("n_images", ctypes.c_int32),
                ("height", ctypes.c_int32),
                ("width", ctypes.c_int32),
                ("spatial_scale", ctypes.c_float)]


class Options(ctypes.Structure):
    """rpool_options: per-call knobs, all zero = defaults."""
    _fields_ = [("cta_threads", ctypes.c_int32),
                ("schedule", ctypes.c_int32),
                ("force_path", ctypes.c_int32),
                ("fuse_heads_backward", ctypes.c_int32),
                ("prefetch_rows", ctypes.c_int32),
                ("prefetch_rois", ctypes.c_int32),
                ("zero_fill_in_tail", ctypes.c_int32),
                ("backward_variant", ctypes.c_int32)]


OPTION_NAMES = tuple(n for n, _ in Options._fields_ if n != "reserved")


class Problem(ctypes.Structure):
    _fields_ = [("n_levels", ctypes.c_int32),
                ("channels", ctypes.c_int32),
                ("feat_layout", ctypes.c_int32),
                ("pool_layout", ctypes.c_int32),
                ("level", Level * MAX_LEVELS),
                ("rois", ctypes.c_void_p),
                ("n_rois", ctypes.c_int32),
                ("roi_format", ctypes.c_int32),
                ("roi_levels", ctypes.c_void_p),
                ("roi_levels_f32", ctypes.c_void_p),
                ("level_thresholds", ctypes.c_float * MAX_LEVELS),
                ("n_thresholds", ctypes.c_int32),
                ("k_min", ctypes.c_int32),
                ("n_heads", ctypes.c_int32),
                ("out_h", ctypes.c_int32 * MAX_HEADS),
                ("out_w", ctypes.c_int32 * MAX_HEADS),
                ("pooled", ctypes.c_void_p * MAX_HEADS),
                ("sampling_ratio", ctypes.c_int32),
                ("coord_mode", ctypes.c_int32),
                ("accumulate", ctypes.c_int32),
                ("deterministic", ctypes.c_int32),
                ("det_workspace", ctypes.c_void_p),
                ("det_workspace_bytes", ctypes.c_size_t),
                ("opt", Options)]


EXPORTS = [
    "rpool_version", "rpool_last_error", "rpool_launch_count", "rpool_build_id",
    "rpool_level_thresholds", "rpool_assign_levels",
    "rpool_workspace_bytes", "rpool_workspace_bytes_ex", "rpool_problem_size", "rpool_plan", "rpool_forward", "rpool_backward",
    "rpool_zero_fill", "rpool_read_plan", "rpool_nchw_to_nhwc", "rpool_nhwc_to_nchw",
    "rpool_backward_det_bytes", "rpool_status_flags",
]
_NOT_INT = ("rpool_last_error", "rpool_launch_count", "rpool_workspace_bytes",
            "rpool_workspace_bytes_ex", "rpool_problem_size", "rpool_build_id")

_lib = None


def lib():
    """The loaded CDLL (built in-tree on first use when nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("RPOOL_B200_LIB") or _build.LIB_PATH   # override: experiments only
    in_tree = path == _build.LIB_PATH
    if in_tree and _build.is_stale():
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(
                "librpool_b200.so is missing or older than its sources and could not be rebuilt (%s); "
                "a stale binary is never loaded and there is no CPU fallback" % e)
    L = ctypes.CDLL(path)
    i32, vp, f32 = ctypes.c_int32, ctypes.c_void_p, ctypes.c_float
    pp = ctypes.POINTER(Problem)
    L.rpool_version.restype = ctypes.c_int
    L.rpool_last_error.restype = ctypes.c_char_p
    L.rpool_launch_count.restype = ctypes.c_uint64
    L.rpool_build_id.restype = ctypes.c_char_p
    L.rpool_level_thresholds.argtypes = [f32, f32, f32, ctypes.c_int, ctypes.c_int,
                                         ctypes.POINTER(f32)]
    L.rpool_assign_levels.argtypes = [vp, i32, i32, i32, ctypes.POINTER(f32), i32, i32, i32,
                                      vp, vp, vp]
    L.rpool_workspace_bytes.argtypes = [i32]
    L.rpool_workspace_bytes.restype = ctypes.c_size_t
    L.rpool_workspace_bytes_ex.argtypes = [i32, i32, i32]
    L.rpool_workspace_bytes_ex.restype = ctypes.c_size_t
    L.rpool_problem_size.restype = ctypes.c_size_t
    for name in ("rpool_plan", "rpool_forward", "rpool_backward"):
        getattr(L, name).argtypes = [pp, vp, ctypes.c_size_t, vp]
    L.rpool_read_plan.argtypes = [vp, i32, vp, vp, vp]
    L.rpool_backward_det_bytes.argtypes = [pp, vp, ctypes.c_size_t, vp, ctypes.POINTER(ctypes.c_size_t)]
    L.rpool_status_flags.argtypes = [vp, i32, vp, ctypes.POINTER(i32)]
    L.rpool_zero_fill.argtypes = [pp, vp]
    for name in ("rpool_nchw_to_nhwc", "rpool_nhwc_to_nchw"):
        getattr(L, name).argtypes = [vp, vp, i32, i32, i32, i32, vp]
    for name in EXPORTS:
        if name not in _NOT_INT:
            getattr(L, name).restype = ctypes.c_int
    if L.rpool_version() != VERSION:
        raise RuntimeError("librpool_b200.so is version %d, this binding needs %d"
                           % (L.rpool_version(), VERSION))
    if in_tree and _build.sources_present():
        have, want = L.rpool_build_id().decode(), _build.source_hash()
        if have != want:
            raise RuntimeError("librpool_b200.so was built from other sources (build id %s, sources %s): "
                               "rebuild with `python chainer-maskrcnn_b200/_build.py --force`" % (have, want))
    if L.rpool_problem_size() != ctypes.sizeof(Problem):
        raise RuntimeError("rpool_problem layout mismatch: library %d bytes, binding %d bytes"
                           % (L.rpool_problem_size(), ctypes.sizeof(Problem)))
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise RpoolError(rc, lib().rpool_last_error().decode("utf-8", "replace"))


def launch_count():
    return int(lib().rpool_launch_count())


def build_id():
    return lib().rpool_build_id().decode()


def make_options(**kw):
    """rpool_options from keyword arguments (unknown names raise)."""
    o = Options()
    for k, v in kw.items():
        if k not in OPTION_NAMES:
            raise ValueError("unknown option %r (known: %s)" % (k, ", ".join(OPTION_NAMES)))
        setattr(o, k, int(v))
    return o


def level_thresholds_libc(s0=224.0, lvl0=4.0, eps=1e-6, k_min=0, k_max=4):
    out = (ctypes.c_float * max(k_max - k_min, 1))()
    check(lib().rpool_level_thresholds(s0, lvl0, eps, k_min, k_max, out))
    return [out[i] for i in range(k_max - k_min)]
