"""FusedStep -- one pooling step (plan + forward + backward) over static buffers.

What a training loop does around the heads, packaged so that the launch overhead
and the gradient zero fill leave the critical path:

  * every buffer (plan workspace, pooled maps, dense gradients) is allocated once;
  * the zero fill of the dense gradients depends on nothing, so it CAN be forked to a
    side stream (``fork_zero_fill=True``: rpool_zero_fill beside plan + forward, then
    rpool_backward with ``accumulate = 1``).  Inside this helper that does not pay and
    is off by default: the forward pass is bound by HBM writes, a fill beside it only
    takes its bandwidth (measured on configs[0..3]: forward launch +25 us, step
    0 .. +3 % slower than the plain sequence once the step is a CUDA graph).  A
    training loop has better places to hide the fill (the heads' own compute between
    the two pooling calls); rpool_zero_fill exists for that;
  * with ``graph=True`` the whole step (fork and join included) is captured once
    into a CUDA graph and replayed: the RoIs, features and upstream gradients are
    read from their device buffers at replay time, so new data is written into the
    same tensors (``step.rois.copy_(...)``) between replays.  Nothing on the path
    synchronises with the host or sizes anything from device data.

The reference has no counterpart: it dispatches two Chainer function calls per RoI
(fpn_roi_mask_head.py:57-63,74-78).
"""
import torch

from . import _engine, _lib


class FusedStep(object):
    """step = FusedStep(features, rois, levels, spatial_scales, out_sizes, sampling_ratio, gys)
    step.run()            # plan + forward + backward on the current stream
    step.outs, step.grads # pooled maps / dense feature gradients (channels-last), reused every run

    features: list of float32 CUDA tensors (N,C,H_l,W_l), channels-last memory;
    rois: (R,5) float32 [image, y1, x1, y2, x2]; levels: None (assigned on the device
    by the reference rule) or an (R,) tensor; gys: one (R,C,P,P) channels-last upstream
    gradient per pooled size (None: forward only)."""

    def __init__(self, features, rois, levels, spatial_scales, out_sizes, sampling_ratio=1,
                 gys=None, graph=True, deterministic=False, fork_zero_fill=False, options=None,
                 fill_in_tail=True):
        self.features = list(features)
        self.rois, self.levels = rois, levels
        self.scales = list(spatial_scales)
        self.sizes = _engine._norm_sizes(out_sizes if isinstance(out_sizes, list) else [out_sizes])
        self.sampling_ratio = int(sampling_ratio)
        self.gys = None if gys is None else list(gys)
        self.deterministic = bool(deterministic)
        self.fill_in_tail = bool(fill_in_tail)
        self.options = dict(options or {})
        dev = rois.device
        for i, f in enumerate(self.features):
            _engine._require_cuda(f, "features[%d]" % i)
            if not f.is_contiguous(memory_format=torch.channels_last) or f.shape[1] % 4:
                raise ValueError("FusedStep needs channels-last features with C % 4 == 0 "
                                 "(static buffers: no conversion or padding inside the step)")
        R, C = rois.shape[0], self.features[0].shape[1]
        self.outs = [torch.empty((R, C, oh, ow), dtype=torch.float32, device=dev,
                                 memory_format=torch.channels_last) for oh, ow in self.sizes]
        self.grads = None
        if self.gys is not None:
            self.grads = [torch.empty_like(f, memory_format=torch.channels_last) for f in self.features]
        coord = _engine.default_coord_mode(self.sampling_ratio)
        n = _lib.lib().rpool_workspace_bytes_ex(R, len(self.sizes), coord)
        self.workspace = torch.empty(n, dtype=torch.uint8, device=dev)
        # the deterministic variant writes every gradient cell itself: nothing to fork
        self.fork = bool(fork_zero_fill) and self.gys is not None and not self.deterministic
        self._side = torch.cuda.Stream(device=dev) if self.fork else None
        self._det_scratch = None
        self.plan = None
        self.graph = None
        if graph:
            self._capture()

    # -- one step on the current stream ------------------------------------
    def _launch(self, marks=None):
        dev = self.rois.device
        with _engine._on(dev):
            cur = torch.cuda.current_stream(dev)
            ev = None
            if marks:
                marks[0].record(cur)
            if self.fork:
                self._side.wait_stream(cur)                      # fork
                with torch.cuda.stream(self._side):
                    _engine.zero_fill(self.grads)
                    ev = self._side.record_event()
            _, self.plan = _engine.forward(self.features, self.rois, self.levels, self.scales, self.sizes,
                                           sampling_ratio=self.sampling_ratio, roi_format=_lib.ROI_YX,
                                           options=self.options, workspace=self.workspace, out=self.outs)
            if marks:
                marks[1].record(cur)
            if self.gys is not None:
                if ev is not None:
                    cur.wait_event(ev)                           # join
                if self.deterministic and self._det_scratch is None:
                    # sized once for the RoIs at hand (+25 %): later steps reuse it without a host
                    # round trip; windows that no longer fit raise RPOOL_FLAG_DET_SCRATCH (status_flags)
                    n = _engine.det_scratch_bytes(self.plan)
                    self._det_scratch = torch.empty(n + n // 4 + 4096, dtype=torch.uint8, device=dev)
                # (no flag read-back inside the step: it would synchronise; see status_flags)
                # the kernel queued before this call is the step's own forward launch, which does not
                # touch the gradients: the zero fill may use its tail
                _engine.backward(self.plan, self.gys, out=self.grads, accumulate=self.fork,
                                 deterministic=self.deterministic, check_flags=False,
                                 det_scratch=self._det_scratch,
                                 fill_in_tail=self.fill_in_tail and not self.fork)
            if marks:
                marks[2].record(cur)

    def _capture(self):
        dev = self.rois.device
        with _engine._on(dev):
            warm = torch.cuda.Stream(device=dev)
            warm.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(warm):
                self._launch()        # shared-memory attributes and allocator state settle before capture
            torch.cuda.current_stream(dev).wait_stream(warm)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._launch()
            self.graph = g

    def run(self, marks=None):
        """One step on the current stream.  ``marks``: three CUDA events recorded before the
        plan, after the forward launch and after the backward launch (measurement only;
        launches from Python instead of replaying the graph)."""
        if self.graph is not None and not marks:
            self.graph.replay()
        else:
            self._launch(marks)
        return self.outs, self.grads
