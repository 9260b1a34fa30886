"""InferenceSession -- the serving-side public API: a static-shape eval forward captured once
in a CUDA graph (~90 kernel launches -> one graph launch) with pinned-host, double-buffered,
stream-overlapped input/output staging.

    sess = InferenceSession(model, batch=32, in_shape=(12, 288, 288))
    y_dev = sess.forward(x_dev)                  # device-resident input
    sess.submit(x_host_pinned); ...; y = sess.collect()   # host buffers, H2D/D2H overlapped with compute

No collective is involved: eval samples are independent, so N GPUs run N sessions on
disjoint batch shards.
"""
from __future__ import annotations

import collections

import torch

from . import _lib


class InferenceSession:
    def __init__(self, model, batch, in_shape, device=None, use_graph=True, slots=2, serving_fusions=True):
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.model = model.to(self.device).eval()
        self.batch, self.in_shape = batch, tuple(in_shape)
        self.use_graph = use_graph
        self.compute = torch.cuda.Stream(self.device)
        self.h2d = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self.static_in = torch.zeros((batch,) + self.in_shape, device=self.device, dtype=torch.float32)
        self.launches_per_forward = 0
        self.graph = None
        # the serving forward may fuse what plain module calls cannot express (OutConv in the last epilogue, model.py);
        # serving_fusions=False captures exactly the reference-API call sequence
        self._fwd = getattr(self.model, "forward_serving", None) if serving_fusions else None
        if self._fwd is None:
            self._fwd = self.model
        self._capture()
        self.out_shape = tuple(self.static_out.shape)
        # staging slots (device side) so H2D of step i+1 and D2H of step i-1 overlap compute of step i
        self.slots = slots
        self.in_stage = [torch.empty_like(self.static_in) for _ in range(slots)]
        self.out_stage = [torch.empty_like(self.static_out) for _ in range(slots)]
        self.out_host = [torch.empty(self.out_shape, dtype=torch.float32, pin_memory=True) for _ in range(slots)]
        self._h2d_done = [torch.cuda.Event() for _ in range(slots)]
        self._in_free = [torch.cuda.Event() for _ in range(slots)]
        self._out_ready = [torch.cuda.Event() for _ in range(slots)]
        self._d2h_done = [torch.cuda.Event() for _ in range(slots)]
        self._pending = collections.deque()
        self._step = 0

    def _capture(self):
        with torch.cuda.device(self.device), torch.no_grad():
            # warm-up on the compute stream: builds the folded-BN / split-weight caches (their small
            # kernels must not be captured) and sizes the allocator
            self.compute.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.compute):
                for _ in range(2):
                    self.static_out = self._fwd(self.static_in)
            self.compute.synchronize()
            n0 = _lib.launch_count()
            if self.use_graph:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=self.compute):
                    self.static_out = self._fwd(self.static_in)
            else:
                with torch.cuda.stream(self.compute):
                    self.static_out = self._fwd(self.static_in)
            self.launches_per_forward = _lib.launch_count() - n0
            self.compute.synchronize()
        # the graph has the addresses of the folded-BN / split-weight tensors baked in: keep them alive with the session.
        # The session is a SNAPSHOT of the weights at construction; after training the model further call refresh().
        from .modules import cached_tensors
        self._keepalive = cached_tensors(self.model)

    def refresh(self):
        """Re-derive the weight caches and re-capture the graph: call after the model's parameters / BatchNorm statistics
        changed (e.g. more training) -- a session is a snapshot of the weights it was built from."""
        from . import ops
        ops.bump_weights_generation()
        self.model.eval()
        self.graph = None
        self._capture()          # static_out may be a new tensor: read it again after refresh()

    # -- device-resident path ---------------------------------------------------------------
    def _run(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            with torch.no_grad():
                self.static_out = self._fwd(self.static_in)

    def forward(self, x_dev):
        """x_dev: (batch, *in_shape) CUDA tensor -> static output tensor (valid until the next call)."""
        with torch.cuda.stream(self.compute):
            self.compute.wait_stream(torch.cuda.current_stream(self.device))
            self.static_in.copy_(x_dev, non_blocking=True)
            self._run()
        torch.cuda.current_stream(self.device).wait_stream(self.compute)
        return self.static_out

    def replay(self):
        """Re-run the captured forward on whatever static_in holds (kernel-only timing)."""
        with torch.cuda.stream(self.compute):
            self._run()

    # -- host-buffer path ----------------------------------------------------------------------
    def submit(self, x_host):
        """Enqueue one batch from PINNED host memory; returns immediately."""
        assert x_host.is_pinned(), "InferenceSession.submit needs pinned host memory for async copies"
        s = self._step % self.slots
        if self._step >= self.slots:          # slot reuse: its previous D2H must have been collected
            assert len(self._pending) < self.slots, "collect() results before submitting more batches"
        with torch.cuda.stream(self.h2d):
            self.h2d.wait_event(self._in_free[s])
            self.in_stage[s].copy_(x_host, non_blocking=True)
            self._h2d_done[s].record(self.h2d)
        with torch.cuda.stream(self.compute):
            self.compute.wait_event(self._h2d_done[s])
            self.static_in.copy_(self.in_stage[s], non_blocking=True)
            self._in_free[s].record(self.compute)
            self._run()
            self.compute.wait_event(self._d2h_done[s])
            self.out_stage[s].copy_(self.static_out, non_blocking=True)
            self._out_ready[s].record(self.compute)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(self._out_ready[s])
            self.out_host[s].copy_(self.out_stage[s], non_blocking=True)
            self._d2h_done[s].record(self.d2h)
        self._pending.append(s)
        self._step += 1

    def collect(self):
        """Block until the oldest submitted batch is back in pinned host memory; returns that tensor
        (reused after ``slots`` further submits)."""
        s = self._pending.popleft()
        self._d2h_done[s].synchronize()
        return self.out_host[s]

    @property
    def h2d_bytes_per_step(self):
        return self.static_in.numel() * 4

    @property
    def d2h_bytes_per_step(self):
        return self.static_out.numel() * 4
