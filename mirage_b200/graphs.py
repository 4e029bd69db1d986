"""CUDA-graph capture of whole steps of the hot path.

A MIRAGE step is hundreds to a thousand short kernel launches (171 for a ViT-L encoder forward, about
1040 for a pretraining step); issued one by one from Python the GPU idles between them.  Every kernel
of ``libmirage_b200.so`` is launched on ``torch.cuda.current_stream()`` with static shapes, allocates
nothing itself and never synchronises, so a whole step can be captured once and replayed.

    step = GraphedCallable(fn)      # fn() -> tensor(s); reads its inputs from fixed tensors
    out = step()                    # first call: warm-up + capture; later calls: one graph launch

Anything that must stay on the host (the reference samples the Dirichlet token split on the CPU,
mirage/model.py:205-207) runs before the replay and writes into the fixed input tensors.
"""
from __future__ import annotations

import torch


class GraphedCallable:
    def __init__(self, fn, warmup: int = 2):
        self.fn = fn
        self.warmup = warmup
        self.graph = None
        self.out = None

    def capture(self):
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self.fn()
        cur.wait_stream(side)
        torch.cuda.synchronize()
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
