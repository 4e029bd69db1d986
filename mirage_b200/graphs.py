"""CUDA-graph capture of whole steps of the hot path.

A MIRAGE step is hundreds to a thousand short kernel launches (171 for a ViT-L encoder forward, about
1040 for a pretraining step); issued one by one from Python the GPU idles between them.  Every kernel
of ``libmirage_b200.so`` is launched on ``torch.cuda.current_stream()`` with static shapes, allocates
nothing itself and never synchronises, so a whole step can be captured once and replayed.

    step = GraphedCallable(fn)      # fn() -> tensor(s); reads its inputs from fixed tensors
    out = step()                    # first call: warm-up + capture; later calls: one graph launch

Anything that must stay on the host (the reference samples the Dirichlet token split on the CPU,
mirage/model.py:205-207) runs before the replay and writes into the fixed input tensors.

Weights: the GEMMs read persistent bf16 twins of the fp32 parameters (functional.bf16_weight).  Their
addresses never change, but their CONTENTS only follow the parameters if something refreshes them:
  * ``refresh_weights=True`` (default whenever autograd is enabled at capture time): every twin is marked
    stale before capture, so the fp32 -> bf16 casts are recorded and each replay re-reads the parameters --
    correct with an optimizer stepped eagerly outside the graph and after ``load_state_dict``;
  * ``refresh_weights=False``: no cast is recorded (inference with frozen weights, or a captured step whose
    optimizer -- optim.FusedAdamW -- rewrites the twins itself).  After replacing weights call ``capture()``
    again, or ``functional.bf16_weight`` on them eagerly (it re-casts into the same buffers).
"""
from __future__ import annotations

import torch


class GraphedCallable:
    def __init__(self, fn, warmup: int = 2, refresh_weights=None):
        self.fn = fn
        self.warmup = warmup
        self.refresh_weights = refresh_weights
        self.graph = None
        self.out = None

    def capture(self):
        from . import functional as Fn
        refresh = torch.is_grad_enabled() if self.refresh_weights is None else self.refresh_weights
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self.fn()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        if refresh:
            Fn.invalidate_weight_cache()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.out = self.fn()
        self.graph = g
        return self

    def __call__(self):
        if self.graph is None:
            self.capture()
        self.graph.replay()
        return self.out
