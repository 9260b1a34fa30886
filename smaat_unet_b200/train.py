"""TrainSession -- the training-side public API: one optimisation step of the reference's Lightning loop
(``training_step`` + ``configure_optimizers``: forward, ``loss_func``, metric update, backward, Adam;
reference models/regression_lightning.py:44-78) as a static-shape step that can be captured in CUDA graphs.

    sess = TrainSession(model, batch=32, in_shape=(12, 288, 288), lr=1e-3)
    loss = sess.step(x, y)          # x: (B, 12, H, W), y: (B, H, W); device or pinned-host tensors; returns a 0-dim device tensor
    sess.metrics.compute()          # PrecipitationMetrics over the steps so far

Everything in a step is enqueued without a host sync: ~1 800 kernel launches (C ABI + Adam's multi-tensor kernels)
collapse into one graph launch (two with data parallelism: the gradient all-reduce runs between them).

Data parallelism (BASELINE configs[3]): one process per GPU, each with its own TrainSession on its shard
of the global batch; gradients live in ONE flat fp32 bucket (4 033 537 floats = 16.1 MB for SmaAt-UNet) that is
all-reduced (average) by NCCL in place -- the parameters' ``.grad`` are views into it, so there is no gather/scatter
copy around the collective.  BatchNorm statistics stay per rank, as in the reference (no SyncBatchNorm).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib, ops
from .metrics import PrecipitationMetrics, step_loss


class TrainSession:
    def __init__(self, model, batch, in_shape, lr=1e-3, device=None, use_graph=True, metrics=None, warmup=3):
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.model = model.to(self.device).train()
        self.batch, self.in_shape = int(batch), tuple(in_shape)
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.use_graph = bool(use_graph)
        if self.world > 1:
            # replicas must start identical (what DDP / Lightning do at construction): rank 0's parameters AND buffers
            with torch.no_grad():
                for t in list(self.model.parameters()) + list(self.model.buffers()):
                    dist.broadcast(t, src=0)
        self.params = [p for p in self.model.parameters() if p.requires_grad]
        # Adam(lr) as in configure_optimizers (regression_lightning.py:47-48); capturable keeps `step` on the device
        self.opt = torch.optim.Adam(self.params, lr=lr, capturable=self.use_graph, foreach=True)
        self.metrics = metrics if metrics is not None else PrecipitationMetrics(device=self.device)
        self.x = torch.zeros((self.batch,) + self.in_shape, device=self.device, dtype=torch.float32)
        self.y = torch.zeros((self.batch,) + self.in_shape[1:], device=self.device, dtype=torch.float32)
        # flat gradient bucket; .grad of every parameter is a view into it
        self.flat_grad = torch.zeros(sum(p.numel() for p in self.params), device=self.device, dtype=torch.float32)
        self._views, off = [], 0
        for p in self.params:
            self._views.append(self.flat_grad[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.loss = torch.zeros((), device=self.device, dtype=torch.float32)
        self.stream = torch.cuda.Stream(self.device)
        self.g_fwd_bwd = self.g_opt = None
        self.launches_per_step = 0
        self._build(warmup)

    # ---- the two halves of a step (split at the collective) ---------------------------------------------
    def _fwd_bwd(self):
        for p in self.params:
            p.grad = None
        pred = self.model(self.x)
        loss = step_loss(pred, self.y, self.metrics)      # loss_func + metrics.update in one pass (metrics.py)
        loss.backward()
        torch._foreach_copy_(self._views, [p.grad for p in self.params])
        self.loss.copy_(loss.detach())

    def _allreduce(self):
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.AVG)

    def _optimise(self):
        for p, v in zip(self.params, self._views):
            p.grad = v
        self.opt.step()

    def _snapshot(self):
        return [t.detach().clone() for t in list(self.model.parameters()) + list(self.model.buffers())]

    def _restore(self, snap):
        with torch.no_grad():
            for t, s in zip(list(self.model.parameters()) + list(self.model.buffers()), snap):
                t.copy_(s)
            for st in self.opt.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
        self.metrics.load_totals(self._metrics_snap)      # warm-up must not disturb totals the caller already holds

    def _build(self, warmup):
        snap = self._snapshot()          # warm-up steps must not change the model the caller handed in
        self._metrics_snap = self.metrics.totals_snapshot()
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            for _ in range(max(1, warmup)):   # builds caches, sizes the allocator, creates Adam's state
                self._fwd_bwd()
                self._allreduce()
                self._optimise()
            self.stream.synchronize()
            n0 = _lib.launch_count()
            if self.use_graph:
                self.g_fwd_bwd = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.g_fwd_bwd, stream=self.stream):
                    self._fwd_bwd()
                self.g_opt = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.g_opt, stream=self.stream, pool=self.g_fwd_bwd.pool()):
                    self._optimise()
            else:
                self._fwd_bwd()
                self._optimise()
            self.launches_per_step = int(_lib.launch_count() - n0)
            self.stream.synchronize()
        self._restore(snap)
        cur.wait_stream(self.stream)

    # ---- public ------------------------------------------------------------------------------------------
    def _stage(self, x, y):
        """Host batch -> device staging slot on the copy stream (overlaps the previous step's compute), then a
        device-to-device copy into the graph's static inputs on the compute stream."""
        if not hasattr(self, "_slots"):
            self.h2d = torch.cuda.Stream(self.device)
            self._slots = [(torch.empty_like(self.x), torch.empty_like(self.y)) for _ in range(2)]
            self._h2d_done = [torch.cuda.Event() for _ in range(2)]
            self._slot_free = [torch.cuda.Event() for _ in range(2)]
            self._n = 0
        i = self._n % 2
        self._n += 1
        sx, sy = self._slots[i]
        with torch.cuda.stream(self.h2d):
            self.h2d.wait_event(self._slot_free[i])
            sx.copy_(x, non_blocking=True)
            sy.copy_(y.reshape(sy.shape), non_blocking=True)
            self._h2d_done[i].record(self.h2d)
        self.stream.wait_event(self._h2d_done[i])
        self.x.copy_(sx, non_blocking=True)
        self.y.copy_(sy, non_blocking=True)
        self._slot_free[i].record(self.stream)

    def last_h2d_event(self):
        """Event marking the end of the most recent host->device batch copy (None before the first host batch).  Hand it to
        ``PinnedBatchLoader.guard`` so the loader does not overwrite a pinned buffer that is still being copied."""
        if not hasattr(self, "_slots") or self._n == 0:
            return None
        return self._h2d_done[(self._n - 1) % 2]

    def load_batch(self, x, y):
        """Copy a batch into the static input buffers (async).  Host tensors (pinned for true overlap) are staged on a
        separate copy stream so the transfer of step i+1 hides behind the compute of step i."""
        if x.device.type == "cpu":
            self._stage(x, y)
        else:
            self.x.copy_(x, non_blocking=True)
            self.y.copy_(y.reshape(self.y.shape), non_blocking=True)

    def step(self, x=None, y=None):
        """One training step on (x, y) (or on the batch already loaded).  Returns the loss (0-dim device tensor, valid in
        stream order; it is overwritten by the next step)."""
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            if x is not None:
                self.load_batch(x, y)
            if self.use_graph:
                self.g_fwd_bwd.replay()
                self._allreduce()
                self.g_opt.replay()
            else:
                self._fwd_bwd()
                self._allreduce()
                self._optimise()
        cur.wait_stream(self.stream)
        ops.bump_weights_generation()    # parameters / running statistics were written by graph replay: no _version bump
        return self.loss
