"""Input adapters with the reference's API (mirage/input_adapters.py) on B200 kernels.

PatchedInputAdapter   Conv2d(k = stride = P) patch embedding + 2-D sin-cos pos-emb  -> [B, N, D]
SemSegInputAdapter    class-embedding lookup -> Conv2d(k = stride = P) + pos-emb     -> [B, N, D]

The stride-P convolution is a GEMM over flattened patches in (c, ph, pw) order.  For the MIRAGE
image modalities (1 channel, 32x32 patches) the A operand is read straight from the fp32 image by a
5-D TMA box (tf32 tensor-core math, no im2col); bias and the positional-embedding row are added in
the GEMM epilogue.  Parameter names / shapes match the reference state_dict:
``pos_emb [1,D,h,w]`` (frozen), ``proj.weight [D,C,P,P]``, ``proj.bias``, ``class_emb.weight``.
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as Fn
from . import ops
from .utils import build_2d_sincos_posemb, pair, trunc_normal_


class _PosEmbMixin:
    """Caches the [N, D] fp32 row table derived from ``pos_emb`` for a given token grid."""

    def _pos_rows(self, nh: int, nw: int, mode: str) -> torch.Tensor:
        p = self.pos_emb
        key = (nh, nw, p._version, p.data_ptr(), p.device)
        cache = self.__dict__.setdefault('_pos_cache', {})
        if cache.get('key') != key:
            with torch.no_grad():
                t = p.detach().float()
                if tuple(t.shape[-2:]) != (nh, nw):
                    # real resize only for non-native grids (reference: input_adapters.py:104, :232)
                    t = F.interpolate(t, size=(nh, nw), mode=mode, align_corners=False)
                cache['rows'] = t[0].flatten(1).t().contiguous()
            cache['key'] = key
        return cache['rows']


class PatchedInputAdapter(nn.Module, _PosEmbMixin):
    """Adapter for spatial inputs (images).  Reference: mirage/input_adapters.py:12-110."""

    def __init__(self, num_channels: int, stride_level: int, patch_size_full: Union[int, Tuple[int, int]],
                 dim_tokens: Optional[int] = None, sincos_pos_emb: bool = True,
                 learnable_pos_emb: bool = False, image_size: Union[int, Tuple[int]] = 224):
        super().__init__()
        self.num_channels = num_channels
        self.stride_level = stride_level
        self.patch_size_full = pair(patch_size_full)
        self.dim_tokens = dim_tokens
        self.sincos_pos_emb = sincos_pos_emb
        self.learnable_pos_emb = learnable_pos_emb
        self.image_size = pair(image_size)
        self.num_patches = (self.image_size[0] // self.patch_size_full[0]) * \
                           (self.image_size[1] // self.patch_size_full[1])
        self.P_H = max(1, self.patch_size_full[0] // stride_level)
        self.P_W = max(1, self.patch_size_full[1] // stride_level)
        if self.dim_tokens is not None:
            self.init(dim_tokens=dim_tokens)

    def init(self, dim_tokens: int = 768):
        self.dim_tokens = dim_tokens
        h = self.image_size[0] // (self.stride_level * self.P_H)
        w = self.image_size[1] // (self.stride_level * self.P_W)
        if self.sincos_pos_emb:
            self.pos_emb = nn.Parameter(build_2d_sincos_posemb(h=h, w=w, embed_dim=self.dim_tokens),
                                        requires_grad=self.learnable_pos_emb)
        else:
            self.pos_emb = nn.Parameter(torch.zeros(1, self.dim_tokens, h, w))
            trunc_normal_(self.pos_emb, std=0.02)
        self.proj = nn.Conv2d(in_channels=self.num_channels, out_channels=self.dim_tokens,
                              kernel_size=(self.P_H, self.P_W), stride=(self.P_H, self.P_W))

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_emb'}

    # -- kernels ---------------------------------------------------------------------------
    def _tma_patch_path(self, x) -> bool:
        _, C, H, W = x.shape
        gh, gw = H // 32, W // 32
        return (C == 1 and self.P_H == 32 and self.P_W == 32 and x.dtype == torch.float32
                and gw > 0 and 128 % gw == 0 and (gh * gw) % 128 == 0)

    def visible_spec(self, x):
        """Description of this modality for Fn.embed_visible (kept tokens only), or None when the fused path does
        not cover the configuration (then MIRAGEModel.forward embeds every token and gathers, as the reference)."""
        if x.dim() != 4 or not x.is_cuda:
            return None
        B, C, H, W = x.shape
        if not (C == 1 and self.P_H == 32 and self.P_W == 32 and x.dtype == torch.float32 and H % 32 == 0
                and W % 32 == 0) or self.pos_emb.requires_grad:
            return None
        nh, nw = H // 32, W // 32
        meta = {'kind': 'patch32', 'count': nh * nw, 'pos': self._pos_rows(nh, nw, 'bicubic')}
        return meta, (x.contiguous(), self.proj.weight, self.proj.bias)

    def write_tokens(self, x, out_buf, row_map):
        """No-autograd fast path: tokens of this modality written into rows
        ``b * stride + offset + t`` of ``out_buf`` ([B * stride, D] fp32)."""
        B, C, H, W = x.shape
        nh, nw = H // self.P_H, W // self.P_W
        pos = self._pos_rows(nh, nw, 'bicubic')
        if self._tma_patch_path(x):
            Fn.patch_tokens_raw(x, self.proj.weight, self.proj.bias, pos, out=out_buf, row_map=row_map)
        else:
            a = self._im2col(x)
            ops.gemm(a, Fn.bf16_weight(self.proj.weight).reshape(self.dim_tokens, -1), m=a.shape[0],
                     n=self.dim_tokens, k=a.shape[1], bias=self.proj.bias.detach(), residual=pos,
                     res_period=nh * nw, out=out_buf, out_row_map=row_map)

    def _im2col(self, x):
        B, C, H, W = x.shape
        nh, nw = H // self.P_H, W // self.P_W
        if x.is_cuda and self.P_W % 8 == 0:
            return ops.patchify_cast(x.contiguous().float(), self.P_H, self.P_W)   # fused permute + bf16 cast
        a = x.reshape(B, C, nh, self.P_H, nw, self.P_W).permute(0, 2, 4, 1, 3, 5)
        return Fn._as_bf16(a.reshape(B * nh * nw, C * self.P_H * self.P_W).float().contiguous())

    def forward(self, x):
        """x: [B, C, H, W] -> tokens [B, N, D] (fp32)."""
        B, C, H, W = x.shape
        assert self.dim_tokens is not None, 'Need to call init(dim_tokens) function first'
        assert (H % self.P_H == 0) and (W % self.P_W == 0), \
            f'Image sizes {H}x{W} must be divisible by patch sizes {self.P_H}x{self.P_W}'
        nh, nw = H // self.P_H, W // self.P_W
        pos = self._pos_rows(nh, nw, 'bicubic')
        if self._tma_patch_path(x):
            tok = Fn.patch_embed32(x, self.proj.weight, self.proj.bias, pos)
        else:
            w2d = self.proj.weight.reshape(self.dim_tokens, -1)
            tok = Fn.linear(self._im2col(x), w2d, self.proj.bias, out_f32=True)
            tok = tok + pos.repeat(B, 1)
        if self.pos_emb.requires_grad:
            # learnable pos-emb: route its gradient through autograd (frozen in every MIRAGE config)
            tok = tok + (self.pos_emb - self.pos_emb.detach())[0].flatten(1).t().repeat(B, 1)
        return tok.reshape(B, nh * nw, self.dim_tokens)


class SemSegInputAdapter(nn.Module, _PosEmbMixin):
    """Adapter for class-index maps (retinal layer maps).  Reference: mirage/input_adapters.py:113-238."""

    def __init__(self, num_classes: int, stride_level: int, patch_size_full: Union[int, Tuple[int, int]],
                 dim_tokens: Optional[int] = None, sincos_pos_emb: int = True, learnable_pos_emb: int = False,
                 image_size: Union[int, Tuple[int]] = 224, dim_class_emb: int = 64,
                 interpolate_class_emb: bool = False, emb_padding_idx: int = None):
        super().__init__()
        self.num_classes = num_classes
        self.stride_level = stride_level
        self.patch_size_full = pair(patch_size_full)
        self.dim_tokens = dim_tokens
        self.sincos_pos_emb = sincos_pos_emb
        self.learnable_pos_emb = learnable_pos_emb
        self.image_size = pair(image_size)
        self.dim_class_emb = dim_class_emb
        self.interpolate_class_emb = interpolate_class_emb
        self.emb_padding_idx = emb_padding_idx
        if self.emb_padding_idx is not None:
            self.num_classes += 1
        self.P_H = max(1, self.patch_size_full[0] // stride_level)
        self.P_W = max(1, self.patch_size_full[1] // stride_level)
        if self.dim_tokens is not None:
            self.init(dim_tokens=dim_tokens)

    def init(self, dim_tokens: int = 768):
        self.dim_tokens = dim_tokens
        h = self.image_size[0] // (self.stride_level * self.P_H)
        w = self.image_size[1] // (self.stride_level * self.P_W)
        if self.sincos_pos_emb:
            self.pos_emb = nn.Parameter(build_2d_sincos_posemb(h=h, w=w, embed_dim=self.dim_tokens),
                                        requires_grad=self.learnable_pos_emb)
        else:
            self.pos_emb = nn.Parameter(torch.zeros(1, self.dim_tokens, h, w))
            trunc_normal_(self.pos_emb, std=0.02)
        self.class_emb = nn.Embedding(num_embeddings=self.num_classes, embedding_dim=self.dim_class_emb,
                                      padding_idx=self.emb_padding_idx)
        trunc_normal_(self.class_emb.weight, std=0.02)
        if self.interpolate_class_emb:
            # reference input_adapters.py:194-200: bilinear DOWN-sampling of the embedded map by the patch size,
            # then a 1x1 convolution (same state_dict keys: proj.1.weight / proj.1.bias)
            self.proj = nn.Sequential(
                nn.Upsample(scale_factor=(1 / self.P_H, 1 / self.P_W), mode='bilinear'),
                nn.Conv2d(in_channels=self.dim_class_emb, out_channels=self.dim_tokens, kernel_size=1, stride=1))
        else:
            self.proj = nn.Conv2d(in_channels=self.dim_class_emb, out_channels=self.dim_tokens,
                                  kernel_size=(self.P_H, self.P_W), stride=(self.P_H, self.P_W))

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_emb', 'class_emb'}

    def _forward_interpolated(self, x):
        """interpolate_class_emb=True (not used by any MIRAGE config): the embedding lookup and the bilinear
        resampling are library calls, the 1x1 projection runs on the GEMM kernel."""
        B, H, W = x.shape
        nh, nw = H // self.P_H, W // self.P_W
        e = self.class_emb(x).permute(0, 3, 1, 2)                      # [B, E, H, W]
        e = self.proj[0](e)                                            # [B, E, nh, nw]
        a = e.permute(0, 2, 3, 1).reshape(B * nh * nw, self.dim_class_emb)
        conv = self.proj[1]
        tok = Fn.linear(Fn.to_bf16(a.contiguous()), conv.weight.reshape(self.dim_tokens, self.dim_class_emb), conv.bias,
                        out_f32=True)
        pos = F.interpolate(self.pos_emb, size=(nh, nw), mode='bilinear')[0].flatten(1).t()
        return (tok + pos.repeat(B, 1)).reshape(B, nh * nw, self.dim_tokens)

    def visible_spec(self, x):
        """See PatchedInputAdapter.visible_spec."""
        if x.dim() != 3 or not x.is_cuda or x.dtype != torch.int64 or self.interpolate_class_emb:
            return None
        B, H, W = x.shape
        if self.P_W % 8 or H % self.P_H or W % self.P_W or self.pos_emb.requires_grad:
            return None
        nh, nw = H // self.P_H, W // self.P_W
        meta = {'kind': 'semseg', 'count': nh * nw, 'pos': self._pos_rows(nh, nw, 'bilinear'),
                'ph': self.P_H, 'pw': self.P_W}
        return meta, (x.contiguous(), self.proj.weight, self.proj.bias, self.class_emb.weight)

    def write_tokens(self, x, out_buf, row_map):
        B, H, W = x.shape
        nh, nw = H // self.P_H, W // self.P_W
        if self.interpolate_class_emb:
            n, stride, off = row_map
            out_buf.view(B, stride, self.dim_tokens)[:, off:off + n].copy_(self._forward_interpolated(x))
            return
        pos = self._pos_rows(nh, nw, 'bilinear')
        a = ops.semseg_patches(x, Fn.bf16_weight(self.class_emb.weight), self.P_H, self.P_W)
        ops.gemm(a, Fn.bf16_weight(self.proj.weight).reshape(self.dim_tokens, -1), m=a.shape[0],
                 n=self.dim_tokens, k=a.shape[1], bias=self.proj.bias.detach(), residual=pos,
                 res_period=nh * nw, out=out_buf, out_row_map=row_map)

    def forward(self, x):
        """x: [B, H, W] int64 class ids -> tokens [B, N, D] (fp32)."""
        B, H, W = x.shape
        assert self.dim_tokens is not None, 'Need to call init(dim_tokens) function first'
        assert (H % self.P_H == 0) and (W % self.P_W == 0), \
            f'Image sizes {H}x{W} must be divisible by patch sizes {self.P_H}x{self.P_W}'
        nh, nw = H // self.P_H, W // self.P_W
        if self.interpolate_class_emb:
            return self._forward_interpolated(x)
        pos = self._pos_rows(nh, nw, 'bilinear')
        tok = Fn.semseg_embed(x, self.class_emb.weight, self.proj.weight, self.proj.bias, pos,
                              self.P_H, self.P_W)
        if self.pos_emb.requires_grad:
            tok = tok + (self.pos_emb - self.pos_emb.detach())[0].flatten(1).t().repeat(B, 1)
        return tok.reshape(B, nh * nw, self.dim_tokens)
