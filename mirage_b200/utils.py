"""Transformer building blocks with the reference's module API (mirage/utils.py), B200 kernels inside.

Same class names, constructor arguments, sub-module / parameter names and forward signatures as
the reference, so its ``state_dict`` loads unchanged.  Tensors crossing these modules follow the
reference's convention ([B, N, D] fp32 in, fp32 out); inside, the fp32 residual stream is kept as a
flat [B*N, D] matrix and every GEMM / attention runs in bf16 on tcgen05 (see functional.py).
"""
from __future__ import annotations

import math
import warnings
from typing import Callable

import torch
import torch.nn as nn

from . import functional as Fn


def pair(t):
    """mirage/utils.py:13-21."""
    if t is None or isinstance(t, tuple):
        return t
    if isinstance(t, list):
        return tuple(t)
    return (t, t)


def build_2d_sincos_posemb(h, w, embed_dim=1024, temperature=10000.):
    """Fixed 2-D sin-cos table [1, D, h, w]; same values and axis convention as mirage/utils.py:24-41
    (host-side, runs once at construction; the table is a frozen parameter of the state_dict)."""
    assert embed_dim % 4 == 0, 'Embed dimension must be divisible by 4 for 2D sin-cos position embedding'
    quarter = embed_dim // 4
    freq = 1.0 / (temperature ** (torch.arange(quarter, dtype=torch.float32) / quarter))
    along_w = torch.arange(w, dtype=torch.float32).repeat_interleave(h)
    along_h = torch.arange(h, dtype=torch.float32).repeat(w)
    a = along_w[:, None] * freq[None, :]
    b = along_h[:, None] * freq[None, :]
    table = torch.cat([a.sin(), a.cos(), b.sin(), b.cos()], dim=1)
    return table.reshape(1, h, w, embed_dim).permute(0, 3, 1, 2).contiguous()


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    """Truncated normal initialiser with the reference's semantics (mirage/utils.py:44-100):
    inverse-CDF sampling on [a, b] (absolute bounds), in place."""
    if (mean < a - 2 * std) or (mean > b + 2 * std):
        warnings.warn("mean is more than 2 std from [a, b] in trunc_normal_", stacklevel=2)
    cdf = lambda v: 0.5 * (1.0 + math.erf(v / math.sqrt(2.0)))
    lo, hi = cdf((a - mean) / std), cdf((b - mean) / std)
    with torch.no_grad():
        tensor.uniform_(2 * lo - 1, 2 * hi - 1).erfinv_().mul_(std * math.sqrt(2.0)).add_(mean)
        tensor.clamp_(min=a, max=b)
    return tensor


class DropPath(nn.Module):
    """Stochastic depth per sample (mirage/utils.py:103-134).  Identity at rate 0 / eval, which is
    every configuration on the benchmarked path; the stochastic branch is plain PyTorch."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0. or not self.training:
            return x
        keep = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        return x.div(keep) * (keep + torch.rand(shape, dtype=x.dtype, device=x.device)).floor_()

    def extra_repr(self) -> str:
        return f'p={self.drop_prob}'


def _flat(x):
    B, N, D = x.shape
    x2 = x.reshape(B * N, D)
    if x2.dtype != torch.float32:
        x2 = x2.float()
    return x2.contiguous(), B, N


class Mlp(nn.Module):
    """fc1 -> GELU(erf) -> fc2 (mirage/utils.py:137-159).  forward: fp32 [B,N,D] -> fp32 [B,N,Do]."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if act_layer is not nn.GELU:
            raise NotImplementedError('the B200 MLP kernel fuses the exact-erf GELU only')
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward_flat(self, x_bf16, residual=None):
        return Fn.mlp(x_bf16, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, residual)

    def forward(self, x):
        x2, B, N = _flat(x)
        y = self.forward_flat(Fn._as_bf16(x2))
        return self.drop(y.float().reshape(B, N, -1))


class Attention(nn.Module):
    """Fused-qkv self attention (mirage/utils.py:162-188)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = attn_drop
        if attn_drop != 0.:
            raise NotImplementedError('attention dropout is 0 in every MIRAGE configuration')
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward_flat(self, x_bf16, B, N, residual=None):
        qkv = Fn.linear(x_bf16, self.qkv.weight, self.qkv.bias)
        o = Fn.self_attention(qkv, B, N, self.num_heads)
        return Fn.linear(o, self.proj.weight, self.proj.bias, residual=residual)

    def forward(self, x):
        x2, B, N = _flat(x)
        y = self.forward_flat(Fn._as_bf16(x2), B, N)
        return self.proj_drop(y.float().reshape(B, N, -1))


class CrossAttention(nn.Module):
    """q from x, k/v from context (mirage/utils.py:191-223)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.attn_drop = attn_drop
        if attn_drop != 0.:
            raise NotImplementedError('attention dropout is 0 in every MIRAGE configuration')
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward_flat(self, x_bf16, ctx_bf16, B, Nq, Nk, out_f32=False):
        q = Fn.linear(x_bf16, self.q.weight, self.q.bias)
        kv = Fn.linear(ctx_bf16, self.kv.weight, self.kv.bias)
        o = Fn.cross_attention(q, kv, B, Nq, Nk, self.num_heads)
        return Fn.linear(o, self.proj.weight, self.proj.bias, out_f32=out_f32)

    def forward(self, x, context):
        x2, B, N = _flat(x)
        c2, _, M = _flat(context)
        y = self.forward_flat(Fn._as_bf16(x2), Fn._as_bf16(c2), B, N, M, out_f32=True)
        return self.proj_drop(y.reshape(B, N, -1))


class Block(nn.Module):
    """Pre-LN transformer block (mirage/utils.py:226-262): one fused autograd node on the B200."""

    def __init__(self, dim: int, num_heads: int, mlp_ratio: float = 4., qkv_bias: bool = False,
                 drop: float = 0., attn_drop: float = 0., drop_path: float = 0.,
                 act_layer: Callable = nn.GELU, norm_layer: Callable = nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.dim = dim
        self.num_heads = num_heads

    def _fusable(self):
        stochastic = self.training and (isinstance(self.drop_path, DropPath) or self.attn.proj_drop.p > 0
                                        or self.mlp.drop.p > 0)
        return (not stochastic) and self.attn.qkv.bias is not None and isinstance(self.norm1, nn.LayerNorm)

    def forward_flat(self, x2, B, N):
        """x2: fp32 [B*N, D] residual stream -> fp32 [B*N, D]."""
        if self._fusable():
            return Fn.transformer_block(
                x2, B, N, self.num_heads, self.norm1.eps,
                self.norm1.weight, self.norm1.bias, self.attn.qkv.weight, self.attn.qkv.bias,
                self.attn.proj.weight, self.attn.proj.bias, self.norm2.weight, self.norm2.bias,
                self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight, self.mlp.fc2.bias)
        # un-fused path (stochastic depth / dropout active, or bias-free qkv): same kernels, torch glue
        h = Fn.layer_norm(x2, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        a = self.attn.proj_drop(self.attn.forward_flat(h, B, N).float())
        x2 = x2 + self.drop_path(a.reshape(B, N, -1)).reshape(B * N, -1)
        h = Fn.layer_norm(x2, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        m = self.mlp.drop(self.mlp.forward_flat(h).float())
        return x2 + self.drop_path(m.reshape(B, N, -1)).reshape(B * N, -1)

    def forward(self, x):
        x2, B, N = _flat(x)
        return self.forward_flat(x2, B, N).reshape(B, N, -1)
