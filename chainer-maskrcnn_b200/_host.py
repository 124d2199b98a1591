"""Host-buffer staging for callers that hold NumPy arrays (the reference's CPU
entry points take NumPy arrays; here they are shipped to the GPU, computed
there, and shipped back -- no arithmetic happens on the host).

Inputs already living in page-locked memory are copied to the device directly;
pageable inputs go through a pinned bounce buffer from torch's caching host
allocator.  Results come back in pinned buffers exposed as NumPy arrays.
"""
import numpy as np
import torch


def device(index=None):
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: chainer_maskrcnn_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device() if index is None else index)


def is_host_array(a):
    return isinstance(a, np.ndarray)


def h2d(a, dtype=np.float32, dev=None):
    """NumPy array -> CUDA tensor of the same shape (async on the current stream)."""
    a = np.ascontiguousarray(a, dtype=dtype)
    t = torch.from_numpy(a)
    if not t.is_pinned():
        bounce = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        bounce.copy_(t)
        t = bounce
    return t.to(dev or device(), non_blocking=True)


def d2h(t):
    """CUDA tensor (any strides) -> NumPy array with the same logical shape,
    backed by pinned memory; synchronises the current stream."""
    host = torch.empty_strided(t.shape, t.stride(), dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy()


def pinned_empty(shape, dtype=torch.float32):
    return torch.empty(shape, dtype=dtype, pin_memory=True)


def d2h_async(t):
    """Starts the copy of a CUDA tensor into a pinned buffer of the same strides on
    the current stream; the caller synchronises that stream before reading
    ``.numpy()`` of the returned tensor."""
    host = torch.empty_strided(t.shape, t.stride(), dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    return host


_COPY_STREAMS = {}


def copy_streams(dev):
    """(upload, download) side streams of a device: the two copy engines run
    concurrently with each other and with the kernels on the caller's stream."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = (torch.cuda.Stream(device=key), torch.cuda.Stream(device=key))
    return _COPY_STREAMS[key]
