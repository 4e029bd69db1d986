"""SpatialOutputAdapter with the reference's API (mirage/output_adapters.py:22-296) on B200 kernels.

Per-task MultiMAE decoder: ``proj_context`` (D -> 256) -> re-insert mask tokens / add task and
positional embeddings / split queries and context (one indexing kernel) -> LayerNorms ->
cross-attention WITHOUT residual -> ``x + mlp(out_norm(x))`` -> ``depth`` self-attention Blocks ->
``out_proj`` whose GEMM epilogue un-patchifies straight into the [B, C, H, W] prediction.

Parameter names / shapes match the reference state_dict (mask_token, pos_emb, task_embeddings.<d>,
proj_context, decoder.{q,kv,proj}, context_norm, query_norm, out_norm, mlp.fc{1,2},
decoder_transformer.<j>.*, out_proj).  Only the MIRAGE pretraining configuration of the adapter is
accelerated (use_xattn=True, task queries); other option combinations raise NotImplementedError.
"""
from __future__ import annotations

from functools import partial
from typing import Callable, Dict, Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as Fn
from .output_adapter_utils import ConvNeXtBlock
from .utils import Block, CrossAttention, Mlp, build_2d_sincos_posemb, pair, trunc_normal_


class SpatialOutputAdapter(nn.Module):
    def __init__(self, num_channels: int, stride_level: int, patch_size_full: Union[int, Tuple[int, int]],
                 dim_tokens_enc: Optional[int] = None, dim_tokens: int = 256, depth: int = 0,
                 learnable_pos_emb: int = False, image_size: Union[int, Tuple[int, int]] = 224,
                 mlp_ratio: int = 4, num_heads: int = 8, qkv_bias: bool = True, drop_rate: float = 0.0,
                 attn_drop_rate: float = 0.0, drop_path_rate: float = 0.0,
                 norm_layer: Callable = partial(nn.LayerNorm, eps=1e-6), use_task_queries: bool = True,
                 task: Optional[str] = None, context_tasks: Optional[list] = None, use_xattn: bool = True):
        super().__init__()
        self.num_channels = num_channels
        self.stride_level = stride_level
        self.patch_size_full = pair(patch_size_full)
        self.dim_tokens_enc = dim_tokens_enc
        self.dim_tokens = dim_tokens
        self.learnable_pos_emb = learnable_pos_emb
        self.image_size = pair(image_size)
        self.use_task_queries = use_task_queries
        self.task = task
        self.use_xattn = use_xattn
        self.num_heads = num_heads
        assert self.patch_size_full is not None and self.image_size is not None
        if not use_xattn:
            raise NotImplementedError('use_xattn=False is not part of the MIRAGE path')

        self.P_H = max(1, self.patch_size_full[0] // stride_level)
        self.P_W = max(1, self.patch_size_full[1] // stride_level)

        self.task_embeddings = None
        if context_tasks is not None:
            self.task_embeddings = nn.ParameterDict(
                {t: nn.Parameter(torch.zeros(1, 1, self.dim_tokens)) for t in context_tasks})
            for emb in self.task_embeddings.values():
                trunc_normal_(emb, std=0.02)

        self.mask_token = nn.Parameter(torch.zeros(1, 1, self.dim_tokens))

        h = self.image_size[0] // (self.stride_level * self.P_H)
        w = self.image_size[1] // (self.stride_level * self.P_W)
        if not self.learnable_pos_emb:
            self.pos_emb = nn.Parameter(build_2d_sincos_posemb(h=h, w=w, embed_dim=self.dim_tokens),
                                        requires_grad=False)
        else:
            raise NotImplementedError('learnable decoder pos-emb is not part of the MIRAGE path')

        self.decoder = CrossAttention(dim=self.dim_tokens, num_heads=num_heads, qkv_bias=qkv_bias,
                                      attn_drop=attn_drop_rate, proj_drop=drop_rate)
        self.context_norm = norm_layer(self.dim_tokens)
        self.query_norm = norm_layer(self.dim_tokens)
        self.out_norm = norm_layer(self.dim_tokens)
        self.mlp = Mlp(in_features=self.dim_tokens, hidden_features=int(self.dim_tokens * mlp_ratio))

        if depth > 0:
            rates = torch.linspace(0, drop_path_rate, depth).tolist()
            self.decoder_transformer = nn.Sequential(*[
                Block(dim=self.dim_tokens, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                      drop=drop_rate, attn_drop=attn_drop_rate, drop_path=rates[i], norm_layer=norm_layer)
                for i in range(depth)])
        else:
            self.decoder_transformer = nn.Identity()

        self.dim_patch = self.num_channels * self.P_H * self.P_W
        self.out_proj = nn.Linear(self.dim_tokens, self.dim_patch)

        if self.dim_tokens_enc is not None:
            self.init(dim_tokens_enc=self.dim_tokens_enc)

    def init(self, dim_tokens_enc: int = 768):
        self.dim_tokens_enc = dim_tokens_enc
        self.proj_context = nn.Linear(self.dim_tokens_enc, self.dim_tokens)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_emb', 'mask_token', 'task_embeddings'}

    # -- embeddings ------------------------------------------------------------------------------
    def _pos_rows(self, size: Tuple[int, int]) -> torch.Tensor:
        pos = self.pos_emb
        if tuple(pos.shape[-2:]) != tuple(size):
            pos = F.interpolate(pos, size=size, mode='bilinear', align_corners=False)
        return pos[0].flatten(1).t()                                  # [N, Dd]

    def generate_context_embeddings(self, input_info, bs: int, size: Tuple[int, int],
                                    device: Optional[torch.device] = None):
        """[bs, N_all, Dd] task + positional embeddings (API parity with output_adapters.py:164-186)."""
        return self._context_emb_table(input_info, size).unsqueeze(0).expand(bs, -1, -1)

    def _context_emb_table(self, input_info, size) -> torch.Tensor:
        """[N_all, Dd] = task_emb[task(p)] + pos_emb[p]  (added in the reference's order)."""
        pos = self._pos_rows(size)
        rows = []
        for task, info in input_info['tasks'].items():
            n = info['num_tokens']
            if self.task_embeddings is not None and task in self.task_embeddings:
                te = self.task_embeddings[task].reshape(1, -1).expand(n, -1)
            else:
                te = pos.new_zeros(n, self.dim_tokens)
            if info['has_posemb']:
                assert n == pos.shape[0], f"# tokens ({n}) != # pos. embeddings ({pos.shape[0]})"
                te = te + pos
            rows.append(te)
        return torch.cat(rows, dim=0)

    def get_queries_and_context(self, context_tokens, input_info, ids_keep, ids_restore):
        """fp32 [B, n_vis+n_glob, Dd] -> (queries [B, N_task, Dd], context [B, n_vis+n_glob, Dd])."""
        if not (self.use_task_queries and self.task in input_info['tasks']):
            raise NotImplementedError('only task queries (decoder task among the encoder inputs) are supported')
        H, W = input_info['tasks'][self.task]['image_size']
        size = (H // (self.stride_level * self.P_H), W // (self.stride_level * self.P_W))
        emb = self._context_emb_table(input_info, size)
        n_glob = input_info.get('num_global_tokens', 0)
        info = input_info['tasks'][self.task]
        return Fn.dec_assemble(context_tokens, self.mask_token.reshape(-1), emb, ids_keep, ids_restore,
                               info['start_idx'], info['end_idx'] - info['start_idx'], n_glob)

    # -- forward -----------------------------------------------------------------------------------
    def forward(self, encoder_tokens: torch.Tensor, input_info: Dict, ids_keep: torch.Tensor,
                ids_restore: torch.Tensor):
        """encoder_tokens fp32 [B, n_vis+n_glob, D] -> prediction fp32 [B, C, H, W]."""
        assert self.dim_tokens_enc is not None, 'Need to call init(dim_tokens_enc) function first'
        H, W = input_info['tasks'][self.task]['image_size']
        N_H = H // (self.stride_level * self.P_H)
        N_W = W // (self.stride_level * self.P_W)
        B, n_ctx, D = encoder_tokens.shape
        Dd = self.dim_tokens

        enc_b = Fn.cached_bf16(encoder_tokens.reshape(B * n_ctx, D))
        ctx = Fn.linear(enc_b, self.proj_context.weight, self.proj_context.bias, out_f32=True)
        q, c = self.get_queries_and_context(ctx.reshape(B, n_ctx, Dd), input_info, ids_keep, ids_restore)
        n_q = q.shape[1]

        qn = Fn.layer_norm(q.reshape(B * n_q, Dd), self.query_norm.weight, self.query_norm.bias,
                           self.query_norm.eps)
        cn = Fn.layer_norm(c.reshape(B * n_ctx, Dd), self.context_norm.weight, self.context_norm.bias,
                           self.context_norm.eps)
        x = self.decoder.forward_flat(qn, cn, B, n_q, n_ctx, out_f32=True)          # no residual (:279)
        h = Fn.layer_norm(x, self.out_norm.weight, self.out_norm.bias, self.out_norm.eps)
        x = self.mlp.forward_flat(h, residual=x)                                    # x + mlp(out_norm(x))
        if not isinstance(self.decoder_transformer, nn.Identity):
            for blk in self.decoder_transformer:
                x = blk.forward_flat(x, B, n_q)
        geom = (self.num_channels, self.P_H, self.P_W, N_H, N_W)
        return Fn.proj_unpatch(Fn.to_bf16(x), self.out_proj.weight, self.out_proj.bias, geom)


# ---------------------------------------------------------------------------------------------
# segmentation heads of the MIRAGELight / seg-tuning caller
# ---------------------------------------------------------------------------------------------
class Adapter(nn.Module):
    """Token selection shared by the segmentation heads (mirage/output_adapters.py:299-322)."""

    def __init__(self, main_tasks: Union[tuple, list] = ('bscan',)):
        super().__init__()
        self.main_tasks = main_tasks

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def adapt_tokens(self, encoder_tokens, input_info):
        parts = [encoder_tokens[:, input_info['tasks'][t]['start_idx']:input_info['tasks'][t]['end_idx']]
                 for t in self.main_tasks]
        return parts[0] if len(parts) == 1 else torch.cat(parts, dim=-1)

    def _grid(self, input_info):
        if self.image_size is None:
            H, W = input_info['tasks'][self.task]['image_size']
        else:
            H, W = self.image_size
        return H, W, H // self.patch_size[0], W // self.patch_size[1]


def _token_matrix(x: torch.Tensor) -> torch.Tensor:
    """[B, N, D] (possibly a slice of the encoder output) -> bf16 [B*N, D] GEMM operand."""
    B, N, D = x.shape
    return Fn.to_bf16(x.reshape(B * N, D).float().contiguous()) if x.dtype != torch.bfloat16 else x.reshape(B * N, D)


class LinearSegAdapter(Adapter):
    """One 1x1 conv on the patch tokens, then bilinear up-sampling (mirage/output_adapters.py:520-575).
    The conv is a [B*N, D] x [D, classes] tcgen05 GEMM on the token matrix (class count padded to a multiple
    of 8 columns inside ``functional.linear_padded``); ``F.interpolate`` stays a library call."""

    def __init__(self, num_classes, main_tasks: Union[tuple, list] = ('bscan',),
                 patch_size: Union[tuple, list] = [16, 16], interpolate_mode: str = 'bilinear',
                 task: Optional[str] = None, image_size: Optional[Tuple[int, int]] = None, **kwargs):
        super().__init__(main_tasks)
        self.patch_size = patch_size
        self.num_classes = num_classes
        self.interpolate_mode = interpolate_mode
        self.task = task
        self.image_size = image_size
        self.final_layer = nn.Identity()

    def init(self, dim_tokens_enc: int = 768):
        self.final_layer = nn.Conv2d(dim_tokens_enc, self.num_classes, 1)

    def forward(self, encoder_tokens: torch.Tensor, input_info: Dict):
        H, W, N_H, N_W = self._grid(input_info)
        x = self.adapt_tokens(encoder_tokens, input_info)
        B, N, D = x.shape
        assert N == N_H * N_W
        w = self.final_layer.weight.reshape(self.num_classes, D)
        if x.is_cuda and D % 64 == 0:
            y = Fn.linear_padded(_token_matrix(x), w, self.final_layer.bias)          # fp32 [B*N, classes]
        else:
            y = F.linear(x.reshape(B * N, D).float(), w, self.final_layer.bias)
        y = y.reshape(B, N_H, N_W, self.num_classes).permute(0, 3, 1, 2)
        return F.interpolate(y, size=(H, W), mode=self.interpolate_mode)


class ConvNeXtAdapter(Adapter):
    """Token projection -> sub-patch feature map -> ConvNeXt blocks -> 1x1 conv -> up-sampling
    (mirage/output_adapters.py:437-517).  ``proj_dec``, the blocks' pointwise layers and ``final_layer`` run as
    tcgen05 GEMMs; activations stay channels-last [B, H', W', C] between blocks."""

    def __init__(self, num_classes, embed_dim: int = 6144, preds_per_patch: int = 16,
                 main_tasks: Union[tuple, list] = ('bscan',), patch_size: list = [16, 16], depth: int = 4,
                 interpolate_mode: str = 'bilinear', task: Optional[str] = None,
                 image_size: Optional[Tuple[int, int]] = None, **kwargs):
        super().__init__(main_tasks)
        self.patch_size = patch_size
        self.embed_dim = embed_dim
        self.preds_per_patch = preds_per_patch
        self.class_dim = embed_dim // preds_per_patch
        self.num_classes = num_classes
        self.interpolate_mode = interpolate_mode
        self.task = task
        self.image_size = image_size
        self.blocks = nn.Sequential(*[ConvNeXtBlock(dim=self.class_dim) for _ in range(depth)])
        self.final_layer = nn.Conv2d(self.class_dim, self.num_classes, 1)
        self.apply(self._init_weights)

    def init(self, dim_tokens_enc: int = 768):
        self.in_channels = dim_tokens_enc * len(self.main_tasks)
        self.proj_dec = nn.Linear(self.in_channels, self.embed_dim)
        self._init_weights(self.proj_dec)

    def forward(self, encoder_tokens: torch.Tensor, input_info: Dict):
        H, W, N_H, N_W = self._grid(input_info)
        x = self.adapt_tokens(encoder_tokens, input_info)
        B, N, D = x.shape
        assert N == N_H * N_W
        s = int(self.preds_per_patch ** 0.5)
        C = self.class_dim
        if x.is_cuda and D % 64 == 0:
            y = Fn.linear(_token_matrix(x), self.proj_dec.weight, self.proj_dec.bias, out_f32=True)
        else:
            y = F.linear(x.reshape(B * N, D).float(), self.proj_dec.weight, self.proj_dec.bias)
        # 'b n (p c) -> b (n p) c' then 'b (nh nw ph pw) c -> b c (nh ph) (nw pw)', kept channels-last
        y = y.reshape(B, N_H, N_W, s, s, C).permute(0, 1, 3, 2, 4, 5).reshape(B, N_H * s, N_W * s, C).contiguous()
        for blk in self.blocks:
            y = blk.forward_nhwc(y)
        Hs, Ws = N_H * s, N_W * s
        w = self.final_layer.weight.reshape(self.num_classes, C)
        if y.is_cuda and C % 64 == 0:
            y = Fn.linear_padded(Fn.to_bf16(y.reshape(B * Hs * Ws, C)), w, self.final_layer.bias)
        else:
            y = F.linear(y.reshape(B * Hs * Ws, C), w, self.final_layer.bias)
        y = y.reshape(B, Hs, Ws, self.num_classes).permute(0, 3, 1, 2)
        return F.interpolate(y, size=(H, W), mode=self.interpolate_mode)
