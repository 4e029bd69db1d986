"""Autograd-aware building blocks of the MIRAGE hot path, composed from the raw kernels in ops.py.

Data conventions (B200-first, not the reference's):
  * residual streams are fp32 2-D tensors [B*N, D] (tokens flattened) -- the LayerNorm kernel reads
    them, the GEMM residual epilogue writes them;
  * everything that feeds a tensor-core GEMM or the attention kernel is bf16;
  * parameters stay fp32 ``nn.Parameter``s with the reference's names and shapes; a bf16 shadow of
    each weight matrix is cached and refreshed when the parameter's version counter moves.

Every ``torch.autograd.Function`` here has a hand-written backward that calls the same C ABI
(dgrad / wgrad GEMMs, LayerNorm backward, attention backward, ...).
"""
from __future__ import annotations

import weakref
from typing import Optional, Sequence

import torch
from torch.autograd import Function

from . import _lib as L
from . import ops

# ---------------------------------------------------------------------------------------------
# bf16 weight shadows
# ---------------------------------------------------------------------------------------------
# Every fp32 weight that feeds a tensor-core GEMM has ONE persistent bf16 twin.  The buffer's address never
# changes for the life of the parameter, so kernels captured in a CUDA graph keep reading the right memory:
#   * whoever modifies the parameter through PyTorch (an eager optimizer, load_state_dict, .copy_) bumps its
#     version counter; the next bf16_weight() call re-casts INTO the same buffer;
#   * optim.FusedAdamW rewrites parameter and twin together in its update kernel (raw pointers: the version
#     counter does not move, the entry stays valid, no cast kernel runs at all);
#   * graphs.GraphedCallable.capture() calls invalidate_weight_cache() first when the captured region does not
#     update the twins itself, so the casts are recorded and every replay refreshes them.
class _Shadow:
    __slots__ = ("ref", "version", "ptr", "buf")

    def __init__(self, ref, version, ptr, buf):
        self.ref, self.version, self.ptr, self.buf = ref, version, ptr, buf


_shadow: dict[int, _Shadow] = {}   # id(param) -> entry
_shadow_epoch = [0]                # bumped whenever a twin is created or dropped (optimizer tables key on it)


def bf16_weight(p: torch.Tensor) -> torch.Tensor:
    """bf16 twin of an fp32 weight (see above).  The entry is tied to the parameter OBJECT (weak reference):
    ids and allocator blocks are recycled once a model is freed, so (id, data_ptr) alone can match a
    different, differently shaped parameter of a later model."""
    if p.dtype == torch.bfloat16:
        return p
    if not isinstance(p, torch.nn.Parameter):
        # a derived tensor (e.g. a zero-padded weight): nothing to key a persistent twin on
        return ops.cast_bf16(p.detach().contiguous())
    key = id(p)
    ent = _shadow.get(key)
    if ent is not None and ent.ref() is p and ent.ptr == p.data_ptr() and ent.buf.shape == p.shape:
        if ent.version != p._version:
            ops.cast_bf16(p.detach().contiguous(), out=ent.buf)
            ent.version = p._version
        return ent.buf
    w = ops.cast_bf16(p.detach().contiguous())

    def _drop(_r, k=key):
        if _shadow.pop(k, None) is not None:
            _shadow_epoch[0] += 1
    _shadow[key] = _Shadow(weakref.ref(p, _drop), p._version, p.data_ptr(), w)
    _shadow_epoch[0] += 1
    return w


def shadow_of(p: torch.Tensor) -> Optional[torch.Tensor]:
    """The persistent bf16 twin of ``p`` if one exists and is current (used by optim.FusedAdamW)."""
    ent = _shadow.get(id(p))
    if ent is None or ent.ref() is not p or ent.ptr != p.data_ptr():
        return None
    if ent.version != p._version:
        ops.cast_bf16(p.detach().contiguous(), out=ent.buf)
        ent.version = p._version
    return ent.buf


def shadow_epoch() -> int:
    return _shadow_epoch[0]


def invalidate_weight_cache():
    """Mark every twin stale: the next use re-casts in place (same address)."""
    for ent in _shadow.values():
        ent.version = -1


def clear_weight_cache():
    _shadow.clear()
    _shadow_epoch[0] += 1


def grad_needed(*tensors) -> bool:
    """True when autograd is recording and some input wants a gradient.  Evaluated OUTSIDE
    Function.forward (inside it grad mode is always off), and passed in as the ``save`` flag."""
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


# bf16 twin of the most recent fp32 gradient a LayerNorm-backward kernel produced: the kernel writes
# both, and the consumer (the next block's backward, or the GEMMs right after) asks for the bf16 view
# through _as_bf16 without a cast kernel re-reading the fp32 tensor.  The fp32 tensor is kept alive
# by the entry, so its storage cannot be recycled for a look-alike while the entry exists.
# The same kernel can also leave the COLUMN SUMS of that gradient (fp32 [D]): for the block below they are the
# bias gradient of its Mlp.fc2, which would otherwise cost a column-sum pass over the whole tensor.
_twin: dict = {"src": None, "ver": -1, "bf16": None, "colsum": None}


def _remember_twin(src: torch.Tensor, bf16: torch.Tensor, colsum: Optional[torch.Tensor] = None):
    _twin["src"], _twin["ver"], _twin["bf16"], _twin["colsum"] = src, src._version, bf16, colsum


def _take_twin(t: torch.Tensor):
    """(bf16 twin, column sums or None) of ``t`` if it is the gradient the last LayerNorm backward produced."""
    src = _twin["src"]
    if (src is not None and t.dtype == torch.float32 and t.data_ptr() == src.data_ptr() and t.shape == src.shape
            and t.stride() == src.stride() and t._version == _twin["ver"]):
        out = (_twin["bf16"], _twin["colsum"])
        _twin["src"] = _twin["bf16"] = _twin["colsum"] = None
        return out
    return None


def _as_bf16(t: torch.Tensor) -> torch.Tensor:
    if t.dtype == torch.bfloat16:
        return t if t.stride(-1) == 1 else t.contiguous()
    hit = _take_twin(t)
    if hit is not None:
        return hit[0]
    return ops.cast_bf16(t.contiguous())


def _sm_count() -> int:
    return int(L.lib().mb_sm_count())   # follows mb_set_sm_reserve


def _wgrad_splits(n_out: int, k_in: int, tokens: int) -> int:
    tiles = ((n_out + 127) // 128) * ((k_in + 255) // 256)
    sms = _sm_count()
    kblocks = (tokens + 63) // 64
    s = max(1, min(sms // max(tiles, 1), kblocks // 4))
    return max(1, s)


# ---------------------------------------------------------------------------------------------
# gradient sink: parameter gradients written straight into their final buffers
# ---------------------------------------------------------------------------------------------
# A data-parallel trainer keeps every ``param.grad`` as a view into a flat all-reduce bucket
# (mirage_b200/ddp.py).  When such a sink is registered, the backward kernels ACCUMULATE into that
# view (wgrad: vector reductions in the GEMM epilogue; bias / LayerNorm grads: accumulate flag) and
# the autograd node returns ``None`` for the parameter, so the engine neither materialises a
# temporary gradient nor launches an add kernel per parameter; the sink is told when the gradient
# is complete (it may then launch the bucket's all-reduce).  Without a sink nothing changes.
_grad_sink = None


def set_grad_sink(sink):
    """sink.target(param) -> fp32 tensor to accumulate into (or None); sink.done(param)."""
    global _grad_sink
    _grad_sink = sink


def _sink_of(param):
    s = _grad_sink
    if s is None or param is None or not param.requires_grad:
        return None
    return s.target(param)


def _done(param, tgt, value):
    """Returns what the autograd node should hand back for ``param``."""
    if tgt is None:
        return value
    _grad_sink.done(param)
    return None


def wgrad(dy: torch.Tensor, x: torch.Tensor, into: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dW[n_out, k_in] = dy[T, n_out]^T x[T, k_in] in fp32 (split-K over tokens with atomics).
    ``into``: accumulate into this [n_out, k_in]-shaped fp32 buffer instead of returning a new one."""
    T, n_out = dy.shape
    k_in = x.shape[1]
    splits = _wgrad_splits(n_out, k_in, T)
    if into is not None:
        out = into.view(n_out, k_in)
        ops.gemm(dy, x, m=n_out, n=k_in, k=T, a_layout=L.MB_MAJOR_MN, b_layout=L.MB_MAJOR_MN, out=out,
                 k_splits=splits, atomic=True)
        return out
    if splits > 1:
        out = torch.zeros((n_out, k_in), dtype=torch.float32, device=dy.device)
    else:
        out = torch.empty((n_out, k_in), dtype=torch.float32, device=dy.device)
    ops.gemm(dy, x, m=n_out, n=k_in, k=T, a_layout=L.MB_MAJOR_MN, b_layout=L.MB_MAJOR_MN, out=out,
             k_splits=splits, atomic=splits > 1)
    return out


def dgrad(dy: torch.Tensor, w_bf16: torch.Tensor, *, out_dtype=torch.bfloat16, dgelu_aux=None,
          colsum_out=None) -> torch.Tensor:
    """dx[T, k_in] = dy[T, n_out] W[n_out, k_in]  (W consumed MN-major: no transposed copy).
    ``colsum_out`` (with ``dgelu_aux``): fp32 [k_in] buffer that ACCUMULATES the column sums of dx --
    the bias gradient of the layer below, computed in the epilogue instead of a separate pass."""
    T, n_out = dy.shape
    k_in = w_bf16.shape[1]
    return ops.gemm(dy, w_bf16, m=T, n=k_in, k=n_out, b_layout=L.MB_MAJOR_MN, out_dtype=out_dtype,
                    dgelu_aux=dgelu_aux, colsum_out=colsum_out)


def _wb_grads(weight, bias, dyb, x, need_w=True, need_b=True):
    """(dW, db) of y = x W^T + b as the autograd node should return them: with a gradient sink set, the wgrad GEMM
    and the column-sum kernel accumulate straight into the parameter's bucket view and None is returned (no
    temporary, no AccumulateGrad add kernel); otherwise fresh tensors."""
    dw = db = None
    if need_w:
        tw = _sink_of(weight)
        dw = _done(weight, tw, wgrad(dyb, x, into=tw))
    if bias is not None and need_b:
        tb = _sink_of(bias)
        db = _done(bias, tb, ops.colsum(dyb, into=tb))
    return dw, db


# ---------------------------------------------------------------------------------------------
# Linear
# ---------------------------------------------------------------------------------------------
class _Linear(Function):
    """y = x W^T + b (+ residual), x bf16 [T, K]; y bf16 or fp32."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, out_f32, save):
        wb = bf16_weight(weight)
        T, K = x.shape
        N = weight.shape[0]
        y = ops.gemm(x, wb, m=T, n=N, k=K, bias=bias, residual=residual,
                     out_dtype=torch.float32 if (out_f32 or residual is not None) else torch.bfloat16)
        if save:
            ctx.save_for_backward(x, weight, bias)
        ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias = ctx.saved_tensors
        dyb = _as_bf16(dy)
        dx = dgrad(dyb, bf16_weight(weight)) if ctx.needs_input_grad[0] else None
        dw, db = _wb_grads(weight, bias, dyb, x, ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        dres = dy if (ctx.has_res and ctx.needs_input_grad[3]) else None
        return dx, dw, db, dres, None, None


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
           residual: Optional[torch.Tensor] = None, out_f32: bool = False) -> torch.Tensor:
    return _Linear.apply(x, weight, bias, residual, out_f32, grad_needed(x, weight, bias, residual))


# Linear with a narrow output (segmentation class logits): the GEMM wants N % 8 == 0, so weight and bias are
# zero-padded to the next multiple of 8 rows (differentiably) and the extra columns are sliced off again.
def linear_padded(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    n = weight.shape[0]
    pad = (-n) % 8
    if pad:
        weight = torch.cat([weight, weight.new_zeros(pad, weight.shape[1])], dim=0)
        if bias is not None:
            bias = torch.cat([bias, bias.new_zeros(pad)], dim=0)
    y = linear(x, weight, bias, out_f32=True)
    return y[:, :n] if pad else y


# ---------------------------------------------------------------------------------------------
# LayerNorm
# ---------------------------------------------------------------------------------------------
class _LayerNorm(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_f32, save):
        if save:
            y, mean, rstd = ops.layernorm(x, weight, bias, eps, save_stats=True,
                                          out_dtype=torch.float32 if out_f32 else torch.bfloat16)
            ctx.save_for_backward(x, weight, bias, mean, rstd)
        else:
            y = ops.layernorm(x, weight, bias, eps, out_dtype=torch.float32 if out_f32 else torch.bfloat16)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias, mean, rstd = ctx.saved_tensors
        tw, tb = _sink_of(weight), _sink_of(bias)
        if tw is None or tb is None:                  # the kernel accumulates both or neither
            tw = tb = None
        dx, dw, db = ops.layernorm_bwd(dy.contiguous(), x, weight, mean, rstd, dw_into=tw, db_into=tb)
        return dx, _done(weight, tw, dw), _done(bias, tb, db), None, None, None


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-6,
               out_f32: bool = False) -> torch.Tensor:
    """x fp32 [T, D] -> bf16 (GEMM operand) or fp32."""
    return _LayerNorm.apply(x, weight, bias, eps, out_f32, grad_needed(x, weight, bias))


# ---------------------------------------------------------------------------------------------
# attention
# ---------------------------------------------------------------------------------------------
class _SelfAttention(Function):
    """qkv bf16 [B*N, 3D] (fused qkv Linear output) -> o bf16 [B*N, D]."""

    @staticmethod
    def forward(ctx, qkv, B, N, heads, need):
        D = qkv.shape[1] // 3
        hd = D // heads
        lse = torch.empty((B, heads, N), dtype=torch.float32, device=qkv.device) if need else None
        o = ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], batch=B, heads=heads, nq=N, nk=N,
                          head_dim=hd, scale=hd ** -0.5, lse=lse)
        if need:
            ctx.save_for_backward(qkv, o, lse)
            ctx.dims = (B, N, heads, hd)
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, o, lse = ctx.saved_tensors
        B, N, heads, hd = ctx.dims
        D = heads * hd
        dqkv = torch.empty_like(qkv)
        ops.attention_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, _as_bf16(do), lse,
                          dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:],
                          batch=B, heads=heads, nq=N, nk=N, head_dim=hd, scale=hd ** -0.5)
        return dqkv, None, None, None, None


class _CrossAttention(Function):
    """q bf16 [B*Nq, D], kv bf16 [B*Nk, 2D] -> o bf16 [B*Nq, D]."""

    @staticmethod
    def forward(ctx, q, kv, B, Nq, Nk, heads, need):
        D = q.shape[1]
        hd = D // heads
        lse = torch.empty((B, heads, Nq), dtype=torch.float32, device=q.device) if need else None
        o = ops.attention(q, kv[:, :D], kv[:, D:], batch=B, heads=heads, nq=Nq, nk=Nk, head_dim=hd,
                          scale=hd ** -0.5, lse=lse)
        if need:
            ctx.save_for_backward(q, kv, o, lse)
            ctx.dims = (B, Nq, Nk, heads, hd)
        return o

    @staticmethod
    def backward(ctx, do):
        q, kv, o, lse = ctx.saved_tensors
        B, Nq, Nk, heads, hd = ctx.dims
        D = heads * hd
        dq = torch.empty_like(q)
        dkv = torch.empty_like(kv)
        ops.attention_bwd(q, kv[:, :D], kv[:, D:], o, _as_bf16(do), lse, dq, dkv[:, :D], dkv[:, D:],
                          batch=B, heads=heads, nq=Nq, nk=Nk, head_dim=hd, scale=hd ** -0.5)
        return dq, dkv, None, None, None, None, None


def self_attention(qkv, B, N, heads):
    return _SelfAttention.apply(qkv, B, N, heads, grad_needed(qkv))


def cross_attention(q, kv, B, Nq, Nk, heads):
    return _CrossAttention.apply(q, kv, B, Nq, Nk, heads, grad_needed(q, kv))


# ---------------------------------------------------------------------------------------------
# MLP:  y = (residual +) fc2(gelu(fc1(x)))
# ---------------------------------------------------------------------------------------------
class _Mlp(Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, residual, need):
        T, D = x.shape
        Hd = w1.shape[0]
        pre = torch.empty((T, Hd), dtype=torch.bfloat16, device=x.device) if need else None
        g = ops.gemm(x, bf16_weight(w1), m=T, n=Hd, k=D, bias=b1, gelu=True, aux_out=pre)
        y = ops.gemm(g, bf16_weight(w2), m=T, n=w2.shape[0], k=Hd, bias=b2, residual=residual,
                     out_dtype=torch.float32 if residual is not None else torch.bfloat16)
        if need:
            ctx.save_for_backward(x, w1, b1, w2, b2, pre, g)
            ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w1, b1, w2, b2, pre, g = ctx.saved_tensors
        dyb = _as_bf16(dy)
        dw2, db2 = _wb_grads(w2, b2, dyb, g)
        dpre = dgrad(dyb, bf16_weight(w2), dgelu_aux=pre)      # (dy W2) * gelu'(pre), bf16 [T, Hd]
        dw1, db1 = _wb_grads(w1, b1, dpre, x)
        dx = dgrad(dpre, bf16_weight(w1)) if ctx.needs_input_grad[0] else None
        dres = dy if ctx.has_res else None
        return dx, dw1, db1, dw2, db2, dres, None


def mlp(x, w1, b1, w2, b2, residual=None):
    return _Mlp.apply(x, w1, b1, w2, b2, residual, grad_needed(x, w1, b1, w2, b2, residual))


# ---------------------------------------------------------------------------------------------
# fused transformer Block (the hot loop): x fp32 [B*N, D] -> fp32 [B*N, D]
#   x1 = x  + proj(attn(qkv(LN1(x))));   x2 = x1 + fc2(gelu(fc1(LN2(x1))))
# mirage/utils.py:259-262.  One autograd node per block: the backward fuses the residual-gradient
# adds into the LayerNorm-backward kernel and never materialises fp32 copies of bf16 gradients.
# ---------------------------------------------------------------------------------------------
class _Block(Function):
    @staticmethod
    def forward(ctx, x, B, N, heads, eps, need, n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b,
                fc1_w, fc1_b, fc2_w, fc2_b):
        T, D = x.shape
        hd = D // heads
        params = (n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b)
        if need:
            h1, mean1, rstd1 = ops.layernorm(x, n1w, n1b, eps, save_stats=True)
        else:
            h1 = ops.layernorm(x, n1w, n1b, eps)
        qkv = ops.gemm(h1, bf16_weight(qkv_w), m=T, n=3 * D, k=D, bias=qkv_b)
        lse = torch.empty((B, heads, N), dtype=torch.float32, device=x.device) if need else None
        a = ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], batch=B, heads=heads, nq=N, nk=N,
                          head_dim=hd, scale=hd ** -0.5, lse=lse)
        x1 = ops.gemm(a, bf16_weight(proj_w), m=T, n=D, k=D, bias=proj_b, residual=x,
                      out_dtype=torch.float32)
        if need:
            h2, mean2, rstd2 = ops.layernorm(x1, n2w, n2b, eps, save_stats=True)
        else:
            h2 = ops.layernorm(x1, n2w, n2b, eps)
        Hd = fc1_w.shape[0]
        pre = torch.empty((T, Hd), dtype=torch.bfloat16, device=x.device) if need else None
        g = ops.gemm(h2, bf16_weight(fc1_w), m=T, n=Hd, k=D, bias=fc1_b, gelu=True, aux_out=pre)
        x2 = ops.gemm(g, bf16_weight(fc2_w), m=T, n=D, k=Hd, bias=fc2_b, residual=x1,
                      out_dtype=torch.float32)
        if need:
            ctx.save_for_backward(x, h1, mean1, rstd1, qkv, a, lse, x1, h2, mean2, rstd2, pre, g, *params)
            ctx.dims = (B, N, heads, hd)
        return x2

    @staticmethod
    def backward(ctx, dx2):
        (x, h1, mean1, rstd1, qkv, a, lse, x1, h2, mean2, rstd2, pre, g,
         n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b) = ctx.saved_tensors
        B, N, heads, hd = ctx.dims
        D = heads * hd
        dx2 = dx2.contiguous()
        hit = _take_twin(dx2)          # produced by the block above: bf16 twin + column sums already there
        dyb, dx2_colsum = hit if hit is not None else (_as_bf16(dx2), None)
        ps = (n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b)
        tg = [_sink_of(q) for q in ps]
        for wi, bi in ((0, 1), (6, 7)):
            # the LayerNorm-backward kernel accumulates both parameter gradients or neither
            if tg[wi] is None or tg[bi] is None:
                tg[wi] = tg[bi] = None
        # MLP branch
        d_fc2_w = wgrad(dyb, g, into=tg[10])
        if dx2_colsum is None:
            d_fc2_b = ops.colsum(dyb, into=tg[11])
        elif tg[11] is not None:
            d_fc2_b = tg[11].add_(dx2_colsum)
        else:
            d_fc2_b = dx2_colsum
        # fc1's bias gradient = column sums of dpre, accumulated by the dgrad epilogue that produces dpre
        d_fc1_b = tg[9] if tg[9] is not None else torch.zeros(fc1_w.shape[0], dtype=torch.float32, device=x.device)
        dpre = dgrad(dyb, bf16_weight(fc2_w), dgelu_aux=pre, colsum_out=d_fc1_b)
        d_fc1_w = wgrad(dpre, h2, into=tg[8])
        dh2 = dgrad(dpre, bf16_weight(fc1_w))
        # the same pass leaves colsum(dx1) = Attention.proj's bias gradient
        d_proj_b = tg[5] if tg[5] is not None else torch.empty(D, dtype=torch.float32, device=x.device)
        dx1, dx1b, d_n2w, d_n2b = ops.layernorm_bwd(dh2, x1, n2w, mean2, rstd2, dres=dx2, want_bf16=True,
                                                    dw_into=tg[6], db_into=tg[7], dx_colsum=d_proj_b,
                                                    dx_colsum_accumulate=tg[5] is not None)
        # attention branch
        d_proj_w = wgrad(dx1b, a, into=tg[4])
        da = dgrad(dx1b, bf16_weight(proj_w))
        dqkv = torch.empty_like(qkv)
        ops.attention_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], a, da, lse,
                          dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:],
                          batch=B, heads=heads, nq=N, nk=N, head_dim=hd, scale=hd ** -0.5)
        d_qkv_w = wgrad(dqkv, h1, into=tg[2])
        d_qkv_b = ops.colsum(dqkv, into=tg[3])
        dh1 = dgrad(dqkv, bf16_weight(qkv_w))
        # colsum(dx) is the fc2 bias gradient of the block BELOW (whose backward runs next and starts from dx)
        dx_sum = torch.empty(D, dtype=torch.float32, device=x.device)
        dx, dxb, d_n1w, d_n1b = ops.layernorm_bwd(dh1, x, n1w, mean1, rstd1, dres=dx1, want_bf16=True,
                                                  dw_into=tg[0], db_into=tg[1], dx_colsum=dx_sum)
        _remember_twin(dx, dxb, dx_sum)  # the previous block's backward starts by casting exactly this tensor
        grads = (d_n1w, d_n1b, d_qkv_w, d_qkv_b, d_proj_w, d_proj_b, d_n2w, d_n2b, d_fc1_w, d_fc1_b,
                 d_fc2_w, d_fc2_b)
        # completion is reported in the order the gradients were produced (output side first), which
        # is the order the all-reduce buckets were laid out in
        out = [None] * 12
        for i in (10, 11, 8, 9, 6, 7, 4, 5, 2, 3, 0, 1):
            out[i] = _done(ps[i], tg[i], grads[i])
        return (dx, None, None, None, None, None, *out)


# ---------------------------------------------------------------------------------------------
# inference path of the Block with the LayerNorms folded into the GEMMs around them
# ---------------------------------------------------------------------------------------------
# LayerNorm(x) W^T + b = rstd * (x (gamma * W)^T) - rstd * mean * colsum(gamma * W) + (b + W beta): the GEMM that
# PRODUCES x (proj / fc2, fp32 residual epilogue) also writes bf16(x) and accumulates the row sums the statistics
# need; the GEMM that CONSUMES LayerNorm(x) (qkv / fc1) reads that bf16 twin against gamma * W and applies mean /
# rstd in its epilogue.  The standalone LayerNorm kernel -- 6.4 % of the cfg-2 step at HBM peak, one 4-byte read
# and one 2-byte write per element -- disappears (47 of 48 per ViT-L forward; the first one has no producer GEMM
# with this epilogue).  Accuracy is that of the unfolded bf16 path (checked with emulated roundings on the ViT-L
# fp32 restatement incl. 30-sigma outlier channels: max-rel 6.8e-3 vs 7.3e-3, identical cosine; tests/test_gpu_lnfold.py).
#
# OFF by default (MB_LN_FOLD=1 or functional.LN_FOLD = True enables it).  Measured A/B inside one gpurun call each
# at cfg 2 (profiles/r02_ab_experiments.txt): with the row statistics accumulated by fp32 atomics +1.0 % -- but then
# identical images at different batch positions stop giving bit-identical tokens; with the deterministic form kept
# here (per-part partial sums written by the producer epilogue, added in slot order by a small kernel) -0.5 %.  The
# standalone kernels say -150 us per block (scripts/perf_lnfold.py); the step does not follow because it is
# power-capped and the LayerNorm kernels are its low-power intervals, in which the clocks recover.
import os as _os
LN_FOLD = _os.environ.get("MB_LN_FOLD", "0") == "1"
_fold_cache: dict = {}
_ln_carry: dict = {"src": None, "ver": -1, "twin": None, "stats": None}


def _folded_weights(w, b, gamma, beta):
    """(bf16(gamma * W), c1 = row sums of that rounded matrix, c2 = b + W beta), cached per parameter versions."""
    key = id(w)
    sig = (w._version, b._version, gamma._version, beta._version, w.data_ptr(), gamma.data_ptr())
    ent = _fold_cache.get(key)
    if ent is not None and ent[0]() is w and ent[1] == sig:
        return ent[2]
    with torch.no_grad():
        wp = (w.detach().float() * gamma.detach().float()[None, :]).to(torch.bfloat16).contiguous()
        c1 = wp.float().sum(dim=1).contiguous()
        c2 = (b.detach().float() + w.detach().float() @ beta.detach().float()).contiguous()
    _fold_cache[key] = (weakref.ref(w, lambda _r, k=key: _fold_cache.pop(k, None)), sig, (wp, c1, c2))
    return wp, c1, c2


def _block_infer_fold(x, B, N, heads, eps, n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b,
                      fc2_w, fc2_b):
    T, D = x.shape
    hd = D // heads
    dev = x.device
    c = _ln_carry
    if (c["src"] is not None and c["src"].data_ptr() == x.data_ptr() and c["src"].shape == x.shape
            and c["ver"] == x._version):
        wq, c1, c2 = _folded_weights(qkv_w, qkv_b, n1w, n1b)
        qkv = ops.gemm(c["twin"], wq, m=T, n=3 * D, k=D, bias=c2, ln_stats=c["stats"], ln_c1=c1, ln_eps=eps)
    else:   # the stream does not come from a folding GEMM (first block): standalone LayerNorm
        h1 = ops.layernorm(x, n1w, n1b, eps)
        qkv = ops.gemm(h1, bf16_weight(qkv_w), m=T, n=3 * D, k=D, bias=qkv_b)
    c["src"] = c["twin"] = c["stats"] = None
    a = ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], batch=B, heads=heads, nq=N, nk=N,
                      head_dim=hd, scale=hd ** -0.5)
    stats = torch.empty((2, T, 1 + ops.gemm_ln_parts(T, D), 2), dtype=torch.float32, device=dev)   # every slot is written
    x1b = torch.empty((T, D), dtype=torch.bfloat16, device=dev)
    x1 = ops.gemm(a, bf16_weight(proj_w), m=T, n=D, k=D, bias=proj_b, residual=x, out_dtype=torch.float32,
                  twin_out=x1b, row_stats=stats[0])
    w1, c1, c2 = _folded_weights(fc1_w, fc1_b, n2w, n2b)
    Hd = fc1_w.shape[0]
    g = ops.gemm(x1b, w1, m=T, n=Hd, k=D, bias=c2, gelu=True, ln_stats=stats[0], ln_c1=c1, ln_eps=eps)
    x2b = torch.empty((T, D), dtype=torch.bfloat16, device=dev)
    x2 = ops.gemm(g, bf16_weight(fc2_w), m=T, n=D, k=Hd, bias=fc2_b, residual=x1, out_dtype=torch.float32,
                  twin_out=x2b, row_stats=stats[1])
    c["src"], c["ver"], c["twin"], c["stats"] = x2, x2._version, x2b, stats[1]
    return x2


def transformer_block(x, B, N, heads, eps, n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b,
                      fc1_w, fc1_b, fc2_w, fc2_b):
    need = grad_needed(x, n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b)
    if LN_FOLD and not need and x.is_cuda and x.dtype == torch.float32 and x.shape[1] % 64 == 0:
        return _block_infer_fold(x.contiguous(), B, N, heads, eps, n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b,
                                 fc1_w, fc1_b, fc2_w, fc2_b)
    return _Block.apply(x, B, N, heads, eps, need, n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b,
                        fc1_w, fc1_b, fc2_w, fc2_b)


# ---------------------------------------------------------------------------------------------
# visible-token selection
# ---------------------------------------------------------------------------------------------
class _TokenGather(Function):
    @staticmethod
    def forward(ctx, tokens, ids_keep, global_tokens):
        ctx.save_for_backward(ids_keep)
        ctx.n_src = tokens.shape[1]
        ctx.n_glob = global_tokens.shape[0]
        return ops.token_gather(tokens.contiguous(), ids_keep.contiguous(), global_tokens.contiguous())

    @staticmethod
    def backward(ctx, dout):
        (ids_keep,) = ctx.saved_tensors
        dsrc, dglob = ops.token_gather_bwd(dout.contiguous(), ids_keep.contiguous(), ctx.n_src, ctx.n_glob)
        return dsrc, None, dglob


def token_gather(tokens, ids_keep, global_tokens):
    """tokens fp32 [B, N_all, D], ids_keep int64 [B, n_keep], global_tokens fp32 [n_glob, D]."""
    return _TokenGather.apply(tokens, ids_keep, global_tokens)


# ---------------------------------------------------------------------------------------------
# patch embedding (PatchedInputAdapter.proj + pos-emb add)
# ---------------------------------------------------------------------------------------------
def patch_tokens_raw(img, weight, bias, pos_rows, out=None, row_map=None):
    """No-autograd kernel call: img fp32 [B,1,H,W] (32x32 patches) -> fp32 token rows.

    tf32 tensor-core GEMM fed straight from the image by a 5-D TMA box (no im2col); the epilogue adds
    the conv bias and the positional-embedding row and can write into a slice of a larger
    [B, N_all(+global), D] token buffer through ``row_map = (period, stride, offset)``.
    """
    Bn, _, H, W = img.shape
    D = weight.shape[0]
    n_tok = (H // 32) * (W // 32)
    if out is None:
        out = torch.empty((Bn * n_tok, D), dtype=torch.float32, device=img.device)
    ops.gemm(img.contiguous(), weight.detach().reshape(D, 1024), m=Bn * n_tok, n=D, k=1024,
             a_layout=L.MB_A_PATCH32, img_hw=(H, W), bias=bias.detach(), residual=pos_rows,
             res_period=n_tok, out=out, out_row_map=row_map)
    return out


def patch_wgrad(img, dtok_bf16):
    """dW[D, 1024] of the 32x32 patch embedding; the bf16 im2col exists only here, never in forward."""
    patches = ops.patchify_cast(img.contiguous().float(), 32, 32)   # one pass: fp32 image -> bf16 [B*N, 1024]
    return wgrad(dtok_bf16, patches)


class _PatchEmbed32(Function):
    @staticmethod
    def forward(ctx, img, weight, bias, pos_rows):
        ctx.save_for_backward(img, weight)
        return patch_tokens_raw(img, weight, bias, pos_rows)

    @staticmethod
    def backward(ctx, dout):
        img, weight = ctx.saved_tensors
        dyb = _as_bf16(dout.contiguous())
        dw = patch_wgrad(img, dyb).reshape(weight.shape)
        return None, dw, ops.colsum(dyb), None


def patch_embed32(img, weight, bias, pos_rows):
    """Autograd version: returns fresh fp32 [B*n_tok, D]."""
    return _PatchEmbed32.apply(img, weight, bias, pos_rows)


# ---------------------------------------------------------------------------------------------
# SemSegInputAdapter: embedding lookup + patch projection (+ pos-emb) -- input_adapters.py:226-236
# ---------------------------------------------------------------------------------------------
class _SemSegEmbed(Function):
    @staticmethod
    def forward(ctx, labels, class_emb, weight, bias, pos_rows, ph, pw, save):
        D = weight.shape[0]
        a = ops.semseg_patches(labels.contiguous(), bf16_weight(class_emb), ph, pw)   # [M, E*ph*pw] bf16
        w2d = bf16_weight(weight).reshape(D, -1)
        n_tok = pos_rows.shape[0]
        out = ops.gemm(a, w2d, m=a.shape[0], n=D, k=a.shape[1], bias=bias, residual=pos_rows,
                       res_period=n_tok, out_dtype=torch.float32)
        if save:
            ctx.save_for_backward(labels, a, class_emb, weight)
            ctx.geom = (ph, pw)
        return out

    @staticmethod
    def backward(ctx, dout):
        labels, a, class_emb, weight = ctx.saved_tensors
        ph, pw = ctx.geom
        D = weight.shape[0]
        dyb = _as_bf16(dout.contiguous())
        dw = wgrad(dyb, a).reshape(weight.shape)
        db = ops.colsum(dyb)
        d_a = dgrad(dyb, bf16_weight(weight).reshape(D, -1))                     # [M, E*ph*pw] bf16
        d_emb = ops.class_emb_grad(labels, d_a, class_emb.shape[0], class_emb.shape[1], ph, pw)
        return None, d_emb, dw, db, None, None, None, None


def semseg_embed(labels, class_emb, weight, bias, pos_rows, ph, pw):
    """labels int64 [B,H,W] -> fp32 [B*n_tok, D]."""
    return _SemSegEmbed.apply(labels, class_emb, weight, bias, pos_rows, ph, pw,
                              grad_needed(class_emb, weight, bias))


# ---------------------------------------------------------------------------------------------
# Visible-token embedding (csrc/visible.cu): input adapters + torch.gather + global tokens of a masked forward
# (mirage/model.py:352-356, :384-391) for the KEPT tokens only.  Returns what
# token_gather(cat(adapter outputs), ids_keep, global_tokens) returns -- fp32 [B * (n_keep + n_glob), D] -- without
# ever embedding the ~87 % of the patches the mask drops, in forward or in backward.
#   meta: one dict per modality, in concatenation order:
#     {'kind': 'patch32', 'count': tokens per sample, 'pos': fp32 [count, D] or None}            tensors: img, W, b
#     {'kind': 'semseg',  'count': ..., 'pos': ..., 'ph': P_H, 'pw': P_W}                        tensors: labels, W, b, E
# ---------------------------------------------------------------------------------------------
class _EmbedVisible(Function):
    @staticmethod
    def forward(ctx, meta, ids_keep, global_tokens, need, *tensors):
        n_mod = len(meta)
        D = global_tokens.shape[-1]
        n_glob = global_tokens.numel() // D
        counts = [m['count'] for m in meta]
        starts = [sum(counts[:i]) for i in range(n_mod)]
        row_src, row_cls = ops.visible_rows(ids_keep, starts, counts, n_glob)
        T = row_cls.numel()
        per = []
        k = 0
        for m in meta:
            n = 4 if m['kind'] == 'semseg' else 3
            per.append(tensors[k:k + n])
            k += n
        tok = ops.embed_rows_init(row_src, row_cls, counts, [t[2].detach() for t in per], [m['pos'] for m in meta],
                                  global_tokens.detach().reshape(n_glob, D).contiguous(), D)
        twins = []
        for i, (m, t) in enumerate(zip(meta, per)):
            if m['kind'] == 'patch32':
                a32, a16 = ops.gather_patches32(t[0], row_src[i], True, need)
                ops.gemm(a32, t[1].detach().reshape(D, 1024), m=T, n=D, k=1024, residual=tok, out=tok)
                twins.append(a16)
            else:
                a = ops.semseg_patches(t[0], bf16_weight(t[3]), m['ph'], m['pw'], row_src=row_src[i])
                ops.gemm(a, bf16_weight(t[1]).reshape(D, -1), m=T, n=D, k=a.shape[1], residual=tok, out=tok)
                twins.append(a)
        if need:
            ctx.meta, ctx.n_glob = meta, n_glob
            ctx.n_tensors = [len(t) for t in per]
            ctx.save_for_backward(row_src, row_cls, global_tokens, *[x for x in twins], *tensors)
        return tok

    @staticmethod
    def backward(ctx, dtok):
        meta, n_glob = ctx.meta, ctx.n_glob
        n_mod = len(meta)
        saved = ctx.saved_tensors
        row_src, row_cls, global_tokens = saved[:3]
        twins = saved[3:3 + n_mod]
        tensors = saved[3 + n_mod:]
        D = global_tokens.shape[-1]
        dtok = dtok.contiguous()
        dyb = _as_bf16(dtok)
        cs = ops.class_colsum(dtok, row_cls, n_mod + n_glob)   # bias gradients + global-token gradient, one pass
        grads = []
        k = 0
        for i, m in enumerate(meta):
            n = ctx.n_tensors[i]
            x, w, b = tensors[k], tensors[k + 1], tensors[k + 2]
            tw = _sink_of(w)
            dw = wgrad(dyb, twins[i], into=tw)
            dw = _done(w, tw, dw if tw is not None else dw.reshape(w.shape))
            tb = _sink_of(b)
            if tb is not None:
                tb.add_(cs[i])
            db = _done(b, tb, cs[i])
            grads += [None, dw, db]
            if m['kind'] == 'semseg':
                emb = tensors[k + 3]
                if emb.requires_grad:
                    d_a = dgrad(dyb, bf16_weight(w).reshape(D, -1))
                    grads.append(ops.class_emb_grad(x, d_a, emb.shape[0], emb.shape[1], m['ph'], m['pw'],
                                                    row_src=row_src[i]))
                else:
                    grads.append(None)
            k += n
        dglob = cs[n_mod:].reshape(global_tokens.shape) if global_tokens.requires_grad else None
        return (None, None, dglob, None, *grads)


def embed_visible(meta, tensors, ids_keep, global_tokens):
    """-> fp32 [B * (n_keep + n_glob), D]; ``tensors``: per modality (x, weight, bias[, class_emb]) flattened."""
    params = [t for t in tensors if t.is_floating_point()]
    return _EmbedVisible.apply(meta, ids_keep.contiguous(), global_tokens, grad_needed(global_tokens, *params),
                               *tensors)


# ---------------------------------------------------------------------------------------------
# fp32 -> bf16 cast as an autograd node (feeds a GEMM from an fp32 residual stream)
# ---------------------------------------------------------------------------------------------
class _ToBf16(Function):
    @staticmethod
    def forward(ctx, x):
        return ops.cast_bf16(x.contiguous())

    @staticmethod
    def backward(ctx, dy):
        return dy.float()


def to_bf16(x):
    return x if x.dtype == torch.bfloat16 else _ToBf16.apply(x)


# ---------------------------------------------------------------------------------------------
# decoder queries / context (output_adapters.py:188-246)
# ---------------------------------------------------------------------------------------------
class _DecAssemble(Function):
    @staticmethod
    def forward(ctx, ctx_tok, mask_token, emb, ids_keep, ids_restore, q_start, n_q, n_glob):
        q, c = ops.dec_assemble(ctx_tok.contiguous(), mask_token.contiguous(), emb.contiguous(),
                                ids_keep.contiguous(), ids_restore.contiguous(), q_start, n_q, n_glob)
        ctx.save_for_backward(ids_keep, ids_restore)
        ctx.geom = (q_start, n_glob)
        return q, c

    @staticmethod
    def backward(ctx, dq, dc):
        ids_keep, ids_restore = ctx.saved_tensors
        q_start, n_glob = ctx.geom
        dctx, demb, dmask = ops.dec_assemble_bwd(dq.contiguous(), dc.contiguous(), ids_keep.contiguous(),
                                                 ids_restore.contiguous(), q_start, n_glob)
        return dctx, dmask, demb, None, None, None, None, None


def dec_assemble(ctx_tok, mask_token, emb, ids_keep, ids_restore, q_start, n_q, n_glob):
    """ctx_tok fp32 [B, n_vis+n_glob, Dd]; mask_token [Dd]; emb [N_all, Dd] -> (queries, context)."""
    return _DecAssemble.apply(ctx_tok, mask_token, emb, ids_keep, ids_restore, q_start, n_q, n_glob)


# ---------------------------------------------------------------------------------------------
# out_proj + un-patchify (output_adapters.py:288-294): tokens bf16 [B*N, Dd] -> image fp32 [B,C,H,W]
# ---------------------------------------------------------------------------------------------
class _ProjUnpatch(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, geom, save):
        C_, ph, pw, gh, gw = geom
        T, K = x.shape
        img = ops.gemm(x, bf16_weight(weight), m=T, n=weight.shape[0], k=K, bias=bias,
                       out_dtype=torch.float32, unpatch=geom)
        if save:
            ctx.save_for_backward(x, weight, bias)
            ctx.geom = geom
        return img

    @staticmethod
    def backward(ctx, dimg):
        x, weight, bias = ctx.saved_tensors
        C_, ph, pw, gh, gw = ctx.geom
        dyb = ops.patchify_cast(dimg.contiguous().float(), ph, pw)          # [T, C*ph*pw] bf16
        dx = dgrad(dyb, bf16_weight(weight)) if ctx.needs_input_grad[0] else None
        dw, db = _wb_grads(weight, bias, dyb, x)
        return dx, dw, db, None, None


def proj_unpatch(x, weight, bias, geom):
    return _ProjUnpatch.apply(x, weight, bias, geom, grad_needed(x, weight, bias))


# ---------------------------------------------------------------------------------------------
# masked criteria (criterion.py)
# ---------------------------------------------------------------------------------------------
class _MaskedMSE(Function):
    @staticmethod
    def forward(ctx, pred, target, mask, scale):
        pred = pred.contiguous()
        target = target.contiguous()
        loss, coef = ops.masked_mse_fwd(pred, target, mask, scale)
        ctx.save_for_backward(pred, target, mask, coef)
        ctx.scale = scale
        return loss

    @staticmethod
    def backward(ctx, g):
        pred, target, mask, coef = ctx.saved_tensors
        return ops.masked_mse_bwd(pred, target, mask, coef, g.contiguous().float(), ctx.scale), None, None, None


class _MaskedCE(Function):
    @staticmethod
    def forward(ctx, logits, target, mask, scale, smoothing):
        logits = logits.contiguous()
        target = target.contiguous()
        loss, coef = ops.masked_ce_fwd(logits, target, mask, scale, smoothing)
        ctx.save_for_backward(logits, target, mask, coef)
        ctx.cfg = (scale, smoothing)
        return loss

    @staticmethod
    def backward(ctx, g):
        logits, target, mask, coef = ctx.saved_tensors
        scale, smoothing = ctx.cfg
        return (ops.masked_ce_bwd(logits, target, mask, coef, g.contiguous().float(), scale, smoothing),
                None, None, None, None)


def masked_mse(pred, target, mask, scale):
    return _MaskedMSE.apply(pred, target, mask, scale)


def masked_ce(logits, target, mask, scale, smoothing=0.0):
    return _MaskedCE.apply(logits, target, mask, scale, smoothing)


# ---------------------------------------------------------------------------------------------
# LayerNorm + token mean-pool (classification tail, mirage_wrapper.py:217-244)
# ---------------------------------------------------------------------------------------------
class _LnMeanPool(Function):
    """x fp32 [B, N, D] -> fp32 [B, len(ranges) * D]: for every (begin, end) row range the mean over those
    tokens of LayerNorm(x), concatenated.  The normalised [B, N, D] tensor is never materialised."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, ranges, need):
        x = x.contiguous()
        B, N, D = x.shape
        pooled = torch.empty((B, len(ranges) * D), dtype=torch.float32, device=x.device)
        saved = []
        for i, (r0, r1) in enumerate(ranges):
            _, xm, mean, rstd = ops.ln_meanpool_fwd(x, gamma, beta, eps, r0, r1, pooled=pooled, col_offset=i * D)
            saved += [xm, mean, rstd]
        if need:
            ctx.save_for_backward(x, gamma, beta, *saved)
            ctx.ranges = ranges
        return pooled

    @staticmethod
    def backward(ctx, dp):
        x, gamma, beta, *saved = ctx.saved_tensors
        D = x.shape[2]
        dp = dp.contiguous().float()
        tg_w, tg_b = _sink_of(gamma), _sink_of(beta)
        acc = tg_w is not None and tg_b is not None
        dx = dg = db = None
        for i, (r0, r1) in enumerate(ctx.ranges):
            xm, mean, rstd = saved[3 * i: 3 * i + 3]
            dx, dg, db = ops.ln_meanpool_bwd(dp, i * D, x, gamma, mean, rstd, xm, r0, r1, dx=dx,
                                             d_gamma=tg_w if acc else dg, d_beta=tg_b if acc else db,
                                             accumulate=acc or i > 0)
        if acc:
            _grad_sink.done(gamma)
            _grad_sink.done(beta)
            dg = db = None
        return dx, dg, db, None, None, None


def ln_meanpool(x, gamma, beta, eps, ranges):
    return _LnMeanPool.apply(x, gamma, beta, eps, tuple(ranges), grad_needed(x, gamma, beta))


# one-entry cache: the three output adapters all project the same encoder tokens (K12 in SURVEY.md).
# The source tensor is held STRONGLY until the next different tensor replaces it (callers pass a temporary
# reshape view that would otherwise die at once and never hit).
_last_cast: dict = {}


def cached_bf16(x: torch.Tensor) -> torch.Tensor:
    key = (x.data_ptr(), x._version, tuple(x.shape), tuple(x.stride()), x.requires_grad, torch.is_grad_enabled())
    # 'src' is alive, so a matching (pointer, version, geometry) cannot be recycled storage: same data
    if _last_cast.get('key') == key and _last_cast.get('src') is not None:
        return _last_cast['out']
    out = to_bf16(x)
    _last_cast['key'] = key
    _last_cast['src'] = x
    _last_cast['out'] = out
    return out


def drop_cast_cache():
    _last_cast.clear()
