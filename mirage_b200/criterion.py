"""Masked reconstruction criteria with the reference's API (mirage/criterion.py), fused on the B200.

``MaskedMSELoss(patch_size, stride=1, norm_pix=False)`` and
``MaskedCrossEntropyLoss(patch_size, stride=1, label_smoothing=0.0)``;
``forward(input, target, mask=None) -> 0-d tensor``.  ``mask`` is the per-token task mask
([B, N] with 1 = masked-out, i.e. the tokens the loss is computed on).

One fused reduction kernel per loss reads prediction and target once; the "empty mask" guard of the
reference (a host sync, criterion.py:36/:103) is evaluated on the device, so an all-zero mask gives
a float 0.0 instead of the reference's int64 ``tensor(0)``.
"""
from __future__ import annotations

from typing import Union

import torch
import torch.nn as nn

from . import functional as Fn


def _scale(patch_size, stride):
    if isinstance(patch_size, int):
        return patch_size // stride
    return patch_size[0] // stride


def _prep_mask(mask, B):
    if mask is None:
        return None
    m = mask.reshape(B, -1)
    if m.dtype != torch.int64:
        m = m.long()
    return m.contiguous()


class MaskedCrossEntropyLoss(nn.Module):
    """Cross-entropy over masked-out patches.  Reference: mirage/criterion.py:11-51."""

    def __init__(self, patch_size: Union[tuple, list], stride: int = 1, label_smoothing: float = 0.0):
        super().__init__()
        self.patch_size = patch_size
        self.stride = stride
        self.scale_factor = _scale(patch_size, stride)
        self.label_smoothing = label_smoothing
        self.epoch = 0

    def forward(self, input, target, mask=None):
        B = input.shape[0]
        return Fn.masked_ce(input.float(), target, _prep_mask(mask, B), self.scale_factor,
                            float(self.label_smoothing))


class SpatialLoss(nn.Module):
    def __init__(self):
        super().__init__()
        self.epoch = 0


class MaskedMSELoss(SpatialLoss):
    """L2 over masked-out patches.  Reference: mirage/criterion.py:70-117."""

    def __init__(self, patch_size: int = 16, stride: int = 1, norm_pix=False):
        super().__init__()
        self.patch_size = patch_size
        self.stride = stride
        self.scale_factor = _scale(patch_size, stride)
        self.norm_pix = norm_pix

    def forward(self, input, target, mask=None):
        if self.norm_pix:
            # per-patch target normalisation (off in every MIRAGE config): plain tensor ops
            p = self.scale_factor
            B, C, H, W = target.shape
            t = target.reshape(B, C, H // p, p, W // p, p).permute(0, 2, 4, 3, 5, 1).reshape(B, -1, p * p * C)
            t = (t - t.mean(dim=-1, keepdim=True)) / torch.sqrt(t.var(dim=-1, keepdim=True) + 1e-6)
            target = t.reshape(B, H // p, W // p, p, p, C).permute(0, 5, 1, 3, 2, 4).reshape(B, C, H, W)
        B = input.shape[0]
        return Fn.masked_mse(input.float(), target.float(), _prep_mask(mask, B), self.scale_factor)
