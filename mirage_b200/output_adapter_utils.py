"""ConvNeXt block of the segmentation adapters (reference: mirage/output_adapter_utils.py:8-46) on the B200
row kernels: depthwise 7x7 conv (the one library call here: cuDNN, channels-last, no layout copies) ->
LayerNorm over channels (csrc/rowops.cu) -> pointwise Linear + GELU -> pointwise Linear (+ layer scale)
+ residual, the two pointwise layers as tcgen05 GEMMs with fused bias / GELU / residual epilogues.

Same parameter names as the reference (dwconv, norm, pwconv1, pwconv2, gamma), so checkpoints load unchanged.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as Fn
from .utils import DropPath


class ConvNeXtBlock(nn.Module):
    def __init__(self, dim, drop_path=0., layer_scale_init_value=0.):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones((dim)),
                                  requires_grad=True) if layer_scale_init_value > 0 else None
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()

    def forward_nhwc(self, x: torch.Tensor) -> torch.Tensor:
        """x fp32 [B, H, W, C] (contiguous) -> same.  Activations stay channels-last between blocks: the
        depthwise conv sees them as a channels_last NCHW view, the row kernels as a [B*H*W, C] matrix."""
        B, H, W, C = x.shape
        y = F.conv2d(x.permute(0, 3, 1, 2), self.dwconv.weight, self.dwconv.bias, padding=3, groups=C)
        y = y.permute(0, 2, 3, 1)
        if not y.is_contiguous():
            y = y.contiguous()
        t = y.reshape(B * H * W, C)
        fusable = (x.is_cuda and C % 128 == 0 and C <= 1024 and self.gamma is None
                   and not (self.training and isinstance(self.drop_path, DropPath)))
        if fusable:
            h = Fn.layer_norm(t, self.norm.weight, self.norm.bias, self.norm.eps)
            out = Fn.mlp(h, self.pwconv1.weight, self.pwconv1.bias, self.pwconv2.weight, self.pwconv2.bias,
                         residual=x.reshape(B * H * W, C))
            return out.reshape(B, H, W, C)
        # layer scale / stochastic depth / odd widths: torch glue around the same arithmetic
        h = F.layer_norm(t, (C,), self.norm.weight, self.norm.bias, self.norm.eps)
        h = self.pwconv2(self.act(self.pwconv1(h)))
        if self.gamma is not None:
            h = self.gamma * h
        return x + self.drop_path(h.reshape(B, H, W, C))

    def forward(self, x):
        """Reference layout: [B, C, H, W] in and out."""
        return self.forward_nhwc(x.permute(0, 2, 3, 1).contiguous().float()).permute(0, 3, 1, 2)
