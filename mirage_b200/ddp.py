"""Data-parallel gradient exchange for MultiMAE pretraining / fine-tuning on one B200 box.

The reference is single-process; this is the new capability the north star asks
for: one process per GPU, replicated weights, per-rank batches and masks, and ONE exchange per
step -- a mean all-reduce of the gradients over NCCL (NVLink 5 / NVSwitch), bucketed in reverse
parameter order and launched from autograd hooks so it overlaps the rest of the backward pass.

    ddp = GradBucketAllReduce(model, bucket_mb=64)
    loss.backward()          # buckets fly as soon as their last gradient is accumulated
    ddp.finish()             # wait for the exchange, gradients now hold the cross-rank mean
    optimizer.step()

Gradients live in flat fp32 buckets (``param.grad`` are views), which also gives the optimizer and
the grad-norm computation contiguous memory.  Works with any backend (``gloo`` on CPU in the tests).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist
from torch import nn


class _Bucket:
    __slots__ = ("flat", "params", "pending", "work")

    def __init__(self, flat, params):
        self.flat = flat
        self.params = params
        self.pending = len(params)
        self.work = None


class GradBucketAllReduce:
    def __init__(self, module: nn.Module, bucket_mb: float = 64.0,
                 process_group: Optional[dist.ProcessGroup] = None, average: bool = True,
                 direct: bool = True, reserve_sms: int = 0):
        self.module = module
        # SMs kept free of persistent compute CTAs while a bucket is in flight (0 = none).  The GEMM / attention
        # / LayerNorm-backward kernels are persistent, one CTA per SM with most of its shared memory: without a
        # reserve an NCCL kernel can only start when some compute kernel ends, and the NEXT compute kernel then
        # finds fewer SMs than CTAs and runs a second, nearly empty wave (measured: 1118 -> 1023 TFLOP/s).
        self.reserve_sms = int(reserve_sms)
        self._reserved = False
        self.enabled = True      # False: skip the exchange (bench.py measures the exposed all-reduce time)
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.average = average
        # ncclAvg exists only in NCCL; other backends (gloo in the tests) sum and divide afterwards
        self._nccl = dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        params = [p for p in module.parameters() if p.requires_grad]
        assert params, "no trainable parameters"
        # gradients become ready roughly in reverse registration order (output side first)
        params = params[::-1]
        cap = int(bucket_mb * 1024 * 1024 / 4)
        # every gradient view starts on a 128-byte boundary inside its bucket: the kernels write them
        # with 16-byte vector stores / vector reductions (a 5-element head bias must not shift the rest)
        self._align = 32

        def padded(n):
            return (n + self._align - 1) // self._align * self._align

        groups: List[List[nn.Parameter]] = [[]]
        size = 0
        for p in params:
            if size + padded(p.numel()) > cap and groups[-1]:
                groups.append([])
                size = 0
            groups[-1].append(p)
            size += padded(p.numel())
        self.buckets: List[_Bucket] = []
        self._defer = False
        self._owner = {}
        self._reported = set()
        self._offset = {}
        for g in groups:
            n = sum(padded(p.numel()) for p in g)
            flat = torch.zeros(n, dtype=torch.float32, device=g[0].device)
            off = 0
            for p in g:
                self._offset[p] = off
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += padded(p.numel())
            b = _Bucket(flat, g)
            self.buckets.append(b)
            for p in g:
                self._owner[p] = b
                p.register_post_accumulate_grad_hook(self._on_grad_ready)
        self.num_params = sum(p.numel() for b in self.buckets for p in b.params)
        # let the backward kernels accumulate straight into the bucket views (no temporary gradient,
        # no per-parameter add kernel); see mirage_b200.functional.set_grad_sink
        if direct and params[0].is_cuda:
            from . import functional as Fn
            Fn.set_grad_sink(self)

    # -- gradient sink protocol ------------------------------------------------------------------
    def _view(self, p: torch.Tensor) -> torch.Tensor:
        b = self._owner[p]
        off = self._offset[p]
        return b.flat[off:off + p.numel()].view_as(p)

    def _check_attached(self, p: torch.Tensor):
        """``p.grad`` must still be the bucket view: ``optimizer.zero_grad()`` (set_to_none=True by default)
        or an assignment to ``p.grad`` detaches it, after which autograd would accumulate into a fresh tensor
        while the all-reduce ships the (zero) bucket -- ranks would diverge silently."""
        b = self._owner[p]
        g = p.grad
        if g is None or g.data_ptr() != b.flat.data_ptr() + 4 * self._offset[p]:
            raise RuntimeError(
                "GradBucketAllReduce: a parameter's .grad is no longer a view into its all-reduce bucket "
                "(use ddp.zero_grad() instead of optimizer.zero_grad(), and never assign p.grad)")

    def target(self, p: torch.Tensor):
        """fp32 buffer the backward kernels accumulate p's gradient into (None: not managed here)."""
        if p not in self._owner:
            return None
        self._check_attached(p)
        if self._owner[p].work is not None:
            raise RuntimeError("GradBucketAllReduce: backward entered while a bucket's all-reduce is still in "
                               "flight (call ddp.finish() first, or use ddp.no_sync() to accumulate)")
        return p.grad

    def done(self, p: torch.Tensor):
        self._on_grad_ready(p)

    # -- hooks -----------------------------------------------------------------------------------
    def _on_grad_ready(self, p: torch.Tensor):
        # reached from the post-accumulate hook and/or from the gradient sink's done(): depending on the
        # PyTorch version the hook also fires for a parameter whose autograd node returned None (2.11
        # does), so every parameter is counted at most once per step
        if p in self._reported:
            return
        self._check_attached(p)
        self._reported.add(p)
        b = self._owner[p]
        if b.work is not None:
            raise RuntimeError("GradBucketAllReduce: a gradient arrived for a bucket whose all-reduce is in flight "
                               "(second backward before finish(); use ddp.no_sync() for gradient accumulation)")
        b.pending -= 1
        if b.pending == 0 and not self._defer:
            self._launch(b)

    def _set_reserve(self, on: bool):
        if self.reserve_sms > 0 and on != self._reserved:
            from . import _lib as L
            L.lib().mb_set_sm_reserve(self.reserve_sms if on else 0)
            self._reserved = on

    def _launch(self, b: _Bucket):
        if self.world > 1 and self.enabled:
            self._set_reserve(True)
            op = dist.ReduceOp.AVG if (self.average and self._nccl) else dist.ReduceOp.SUM
            b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)

    # -- gradient accumulation ---------------------------------------------------------------------
    def no_sync(self):
        """Context manager: backward passes inside it only ACCUMULATE into the buckets (no all-reduce is
        launched); the exchange happens in the first backward outside it / in ``finish()``.  Mirrors
        ``DistributedDataParallel.no_sync``."""
        ddp = self

        class _NoSync:
            def __enter__(self_):
                self_.prev = ddp._defer
                ddp._defer = True

            def __exit__(self_, *a):
                ddp._defer = self_.prev
                # hooks fire once per backward: forget who reported so the next backward counts again
                ddp._reported.clear()
                for b in ddp.buckets:
                    b.pending = len(b.params)
                return False
        return _NoSync()

    # -- step protocol -----------------------------------------------------------------------------
    def zero_grad(self, memset: bool = True):
        """Zero the buckets in place (keeps ``param.grad`` as views; never set grads to None).
        ``memset=False`` only resets the bookkeeping: for optimizers that zero the gradients inside their own
        update pass (``optim.FusedAdamW(zero_grad_in_step=True)``)."""
        self._reported.clear()
        for b in self.buckets:
            if memset:
                b.flat.zero_()
            b.pending = len(b.params)
            b.work = None
            for p in b.params:
                off = self._offset[p]
                if p.grad is None or p.grad.data_ptr() != b.flat.data_ptr() + 4 * off:
                    p.grad = b.flat[off:off + p.numel()].view_as(p)

    def finish(self):
        """Block the current stream on every outstanding bucket; flush buckets whose hooks did not
        all fire (parameters unused this step keep a zero gradient, as in DDP)."""
        for b in self.buckets:
            if b.work is None and self.world > 1:
                self._launch(b)
        for b in self.buckets:
            if b.work is not None:
                b.work.wait()
                if self.average and not self._nccl:
                    b.flat.div_(self.world)
                b.work = None
        self._set_reserve(False)

    def grad_norm(self) -> torch.Tensor:
        """Global L2 norm of the (already exchanged) gradients, identical on every rank
        (mutils/native_scaler.py:46-61 computes the same quantity per parameter)."""
        return torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(b.flat) for b in self.buckets]))
