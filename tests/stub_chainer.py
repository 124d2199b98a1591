"""Minimal stand-ins for `chainer` and `cupy` (neither is installed in this image) so that
chainer_maskrcnn_b200.chainer_adapter -- which is written against the real packages'
public API -- can be exercised: a CuPy-like ndarray over torch CUDA memory
(`.data.ptr`, `.shape`, `.dtype`, `.size`, fancy column indexing), `cupy.empty/zeros/
asarray/asnumpy/ascontiguousarray`, `cupy.cuda.get_current_stream().ptr`, and the
old-style `chainer.function.Function` protocol (check_type_forward -> forward_gpu,
retain_inputs, 1-tuple outputs, backward_gpu(inputs, gy)).  TEST INFRASTRUCTURE ONLY.
"""
import sys
import types

import numpy as np
import torch


class _Mem(object):
    def __init__(self, t):
        self.ptr = t.data_ptr()


class ndarray(object):
    def __init__(self, t):
        self._t = t
        self.data = _Mem(t)

    shape = property(lambda self: tuple(self._t.shape))
    ndim = property(lambda self: self._t.dim())
    size = property(lambda self: self._t.numel())
    dtype = property(lambda self: np.dtype(str(self._t.dtype).replace("torch.", "")))

    def __getitem__(self, idx):
        return ndarray(self._t[idx].contiguous())


def _dt(dtype):
    return getattr(torch, np.dtype(dtype).name)


def install():
    """Plants the stand-ins in sys.modules; returns a function that removes them again."""
    cupy = types.ModuleType("cupy")
    cupy.ndarray = ndarray
    cupy.empty = lambda shape, dtype=np.float32: ndarray(torch.empty(tuple(shape), dtype=_dt(dtype), device="cuda"))
    cupy.zeros = lambda shape, dtype=np.float32: ndarray(torch.zeros(tuple(shape), dtype=_dt(dtype), device="cuda"))
    cupy.asarray = lambda a: a if isinstance(a, ndarray) else ndarray(torch.from_numpy(np.ascontiguousarray(a)).cuda())
    cupy.asnumpy = lambda a: a._t.cpu().numpy()
    cupy.ascontiguousarray = lambda a: ndarray(a._t.contiguous())
    cuda = types.ModuleType("cupy.cuda")
    cuda.get_current_stream = lambda: types.SimpleNamespace(ptr=torch.cuda.current_stream().cuda_stream)
    cupy.cuda = cuda

    chainer = types.ModuleType("chainer")
    function = types.ModuleType("chainer.function")
    utils = types.ModuleType("chainer.utils")
    type_check = types.ModuleType("chainer.utils.type_check")

    class InvalidType(Exception):
        pass

    class _Types(list):
        def size(self):
            return len(self)

    def expect(*conds):
        if not all(bool(c) for c in conds):
            raise InvalidType("type check failed")

    type_check.InvalidType, type_check.expect = InvalidType, expect

    class Function(object):
        """forward(inputs) -> tuple; the call protocol of chainer's old-style Function."""

        def retain_inputs(self, indexes):
            self._retained = tuple(indexes)

        def __call__(self, *inputs):
            self.check_type_forward(_Types(inputs))
            self._inputs = inputs
            outs = self.forward_gpu(inputs)
            assert isinstance(outs, tuple)
            return outs[0] if len(outs) == 1 else outs

        def check_type_forward(self, in_types):
            pass

        def backward_from(self, *gys):
            kept = getattr(self, "_retained", tuple(range(len(self._inputs))))
            inputs = tuple(x if i in kept else None for i, x in enumerate(self._inputs))
            return self.backward_gpu(inputs, gys)

    function.Function = Function
    chainer.function, chainer.utils, utils.type_check = function, utils, type_check
    mods = {"cupy": cupy, "cupy.cuda": cuda, "chainer": chainer, "chainer.function": function,
            "chainer.utils": utils, "chainer.utils.type_check": type_check}
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    sys.modules.pop("chainer_maskrcnn_b200.chainer_adapter", None)

    def remove():
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        sys.modules.pop("chainer_maskrcnn_b200.chainer_adapter", None)
    return remove
