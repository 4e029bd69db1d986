"""Thin tensor-level wrappers over the C ABI (raw kernels, no autograd).

Every function enqueues work on ``torch.cuda.current_stream()`` and returns the output tensor(s).
Tensors must live on a CUDA device; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# ---------------------------------------------------------------------------------------------
# launch accounting / per-kernel timing hooks (used by bench.py; zero cost when disabled)
# ---------------------------------------------------------------------------------------------
class _Stats:
    launches = 0          # kernels launched through this module since reset
    recorder = None       # callable(name, work, unit) -> context manager, or None


def reset_launch_count():
    _Stats.launches = 0


def launch_count() -> int:
    return _Stats.launches


def set_recorder(fn):
    """fn(name: str, work: float, unit: 'flop'|'byte') must return a context manager that brackets
    the launch (bench.py records CUDA events on the current stream inside it)."""
    _Stats.recorder = fn


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NULL = _NullCtx()


def _rec(name: str, work: float, unit: str, kernels: int = 1):
    _Stats.launches += kernels
    r = _Stats.recorder
    return _NULL if r is None else r(name, work, unit)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.MirageB200Error("mirage_b200 kernels need CUDA tensors (no CPU fallback)")


def gemm(a: torch.Tensor, b: torch.Tensor, *, m: int, n: int, k: int,
         a_layout: int = L.MB_MAJOR_K, b_layout: int = L.MB_MAJOR_K,
         out: torch.Tensor | None = None, out_dtype: torch.dtype = torch.bfloat16,
         bias: torch.Tensor | None = None,
         residual: torch.Tensor | None = None, res_period: int = 0,
         gelu: bool = False, aux_out: torch.Tensor | None = None,
         dgelu_aux: torch.Tensor | None = None,
         atomic: bool = False, k_splits: int = 1, block_n: int = 0,
         img_hw: tuple[int, int] | None = None,
         out_row_map: tuple[int, int, int] | None = None,
         unpatch: tuple[int, int, int, int, int] | None = None, cta_pair: int = 0,
         colsum_out: torch.Tensor | None = None,
         twin_out: torch.Tensor | None = None, row_stats: torch.Tensor | None = None,
         ln_stats: torch.Tensor | None = None, ln_c1: torch.Tensor | None = None,
         ln_eps: float = 1e-6) -> torch.Tensor:
    """out[m, n] = epilogue(A * B^T); see ``mb_gemm`` in include/mirage_b200.h for the contract.

    ``a`` / ``b`` are 2-D (or, for MB_A_PATCH32, the [B,1,H,W] fp32 image batch) with unit stride in
    the last dimension; leading dimensions are taken from ``stride(0)``.
    """
    _req_cuda(a, b, out, bias, residual, aux_out, dgelu_aux)
    tf32 = a.dtype == torch.float32
    if a.dtype not in (torch.bfloat16, torch.float32) or b.dtype != a.dtype:
        raise L.MirageB200Error(f"gemm: unsupported operand dtypes {a.dtype}, {b.dtype}")
    if out is None:
        if unpatch is not None:
            c_, ph_, pw_, gh_, gw_ = unpatch
            out = torch.empty((m // (gh_ * gw_), c_, gh_ * ph_, gw_ * pw_), dtype=out_dtype, device=a.device)
        else:
            out = torch.empty((m, n), dtype=out_dtype, device=a.device)
    args = L.GemmArgs()
    args.a, args.b, args.out = a.data_ptr(), b.data_ptr(), out.data_ptr()
    args.bias = _ptr(bias)
    args.residual = _ptr(residual)
    args.aux_in = _ptr(dgelu_aux)
    args.aux_out = _ptr(aux_out)
    args.m, args.n, args.k = m, n, k
    if a_layout == L.MB_A_PATCH32:
        assert img_hw is not None and a.is_contiguous()
        args.lda = 0
        args.img_h, args.img_w = img_hw
    else:
        assert a.stride(-1) == 1
        args.lda = a.stride(0)
    assert b.stride(-1) == 1 and out.stride(-1) == 1
    args.ldb = b.stride(0)
    args.ldc = out.stride(0) if unpatch is None else n
    args.ld_res = residual.stride(0) if residual is not None else 0
    aux = aux_out if aux_out is not None else dgelu_aux
    args.ld_aux = aux.stride(0) if aux is not None else 0
    args.res_period = res_period
    args.a_layout, args.b_layout = a_layout, b_layout
    args.in_dtype = L.MB_F32 if tf32 else L.MB_BF16
    args.out_dtype = L.MB_F32 if out.dtype == torch.float32 else L.MB_BF16
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
    if residual is not None:
        assert residual.dtype == torch.float32
    epi = 0
    if gelu:
        epi |= L.MB_EPI_GELU
    if dgelu_aux is not None:
        epi |= L.MB_EPI_DGELU
    if atomic or k_splits > 1:
        epi |= L.MB_EPI_ATOMIC
    args.epilogue = epi
    args.k_splits = k_splits
    args.block_n = block_n
    args.cta_pair = cta_pair
    args.colsum_out = _ptr(colsum_out)
    if twin_out is not None:      # producer of a folded LayerNorm: bf16 twin of the output + row statistics
        assert row_stats is not None and twin_out.dtype == torch.bfloat16 and twin_out.stride(-1) == 1
        assert row_stats.dtype == torch.float32 and row_stats.is_contiguous() and row_stats.dim() == 3
        assert row_stats.shape[0] == m and row_stats.shape[2] == 2     # [m, 1 + gemm_ln_parts(m, n), 2]
        args.twin_out, args.ld_twin, args.row_stats = twin_out.data_ptr(), twin_out.stride(0), row_stats.data_ptr()
        args.ln_parts = row_stats.shape[1] - 1
    if ln_stats is not None:      # consumer: LayerNorm applied in the epilogue
        assert ln_c1 is not None and ln_stats.dtype == torch.float32 and ln_stats.is_contiguous()
        assert ln_stats.dim() == 3 and ln_stats.shape[0] == m and ln_stats.shape[2] == 2
        assert ln_c1.dtype == torch.float32 and ln_c1.is_contiguous() and bias is not None
        args.ln_stats, args.ln_c1, args.ln_eps = ln_stats.data_ptr(), ln_c1.data_ptr(), float(ln_eps)
        args.ln_parts = ln_stats.shape[1] - 1
    if out_row_map is not None:
        args.out_row_period, args.out_row_stride, args.out_row_offset = out_row_map
    if unpatch is not None:
        assert out.is_contiguous()
        args.epilogue |= L.MB_EPI_UNPATCH
        args.up_channels, args.up_ph, args.up_pw, args.up_gh, args.up_gw = unpatch
    with _rec("gemm", 2.0 * m * n * k, "flop"):
        L.check(L.lib().mb_gemm(C.byref(args), _stream()), "mb_gemm")
    return out


def gemm_ln_parts(m: int, n: int) -> int:
    """Partial sums per row of the folded-LayerNorm statistics for an [m, n] producer GEMM (see mb_gemm_ln_parts)."""
    return int(L.lib().mb_gemm_ln_parts(m, n))


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, batch: int, heads: int,
              nq: int, nk: int, head_dim: int, scale: float,
              out: torch.Tensor | None = None, lse: torch.Tensor | None = None) -> torch.Tensor:
    """softmax(q k^T * scale) v.  q/k/v are 2-D bf16 views [B*n, ld] (possibly column slices of one
    fused qkv buffer); returns bf16 [B*nq, heads*head_dim] in token-major / head-minor layout."""
    _req_cuda(q, k, v, out, lse)
    assert q.dtype == k.dtype == v.dtype == torch.bfloat16
    assert q.stride(-1) == 1 and k.stride(-1) == 1 and v.stride(-1) == 1
    if out is None:
        out = torch.empty((batch * nq, heads * head_dim), dtype=torch.bfloat16, device=q.device)
    a = L.AttnArgs()
    a.q, a.k, a.v, a.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.lse = _ptr(lse)
    a.batch, a.heads, a.nq, a.nk = batch, heads, nq, nk
    a.ldq, a.ldk, a.ldv, a.ldo = q.stride(0), k.stride(0), v.stride(0), out.stride(0)
    a.head_dim = head_dim
    a.scale = scale
    with _rec("attn_fwd", 4.0 * batch * heads * nq * nk * head_dim, "flop"):
        L.check(L.lib().mb_attn_fwd(C.byref(a), _stream()), "mb_attn_fwd")
    return out


def layernorm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-6, *,
              out_dtype: torch.dtype = torch.bfloat16, save_stats: bool = False):
    """Row LayerNorm of an fp32 [rows, dim] tensor.  Returns y (and mean, rstd when save_stats)."""
    _req_cuda(x, weight, bias)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    rows, dim = x.shape
    y = torch.empty((rows, dim), dtype=out_dtype, device=x.device)
    mean = rstd = None
    if save_stats:
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    with _rec("layernorm_fwd", rows * dim * (4.0 + y.element_size()), "byte"):
        L.check(L.lib().mb_layernorm_fwd(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), y.data_ptr(),
                                         L.MB_BF16 if out_dtype == torch.bfloat16 else L.MB_F32,
                                         _ptr(mean), _ptr(rstd), rows, dim, x.stride(0), y.stride(0),
                                         eps, _stream()), "mb_layernorm_fwd")
    return (y, mean, rstd) if save_stats else y


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, weight: torch.Tensor, mean: torch.Tensor,
                  rstd: torch.Tensor, dres: torch.Tensor | None = None, want_bf16: bool = False,
                  dw_into: torch.Tensor | None = None, db_into: torch.Tensor | None = None,
                  dx_colsum: torch.Tensor | None = None, dx_colsum_accumulate: bool = False):
    """Returns (dx f32, dweight f32, dbias f32); dx includes dres when given.  With ``want_bf16`` the
    kernel also writes dx rounded to bf16 and (dx, dx_bf16, dweight, dbias) is returned.
    ``dx_colsum``: fp32 [dim] buffer that receives (or, with ``dx_colsum_accumulate``, accumulates) the column
    sums of dx -- the bias gradient of the Linear whose output gradient dx is."""
    _req_cuda(dy, x, weight, mean, rstd, dres)
    rows, dim = x.shape
    dx = torch.empty((rows, dim), dtype=torch.float32, device=x.device)
    dxb = torch.empty((rows, dim), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    # dw_into / db_into: ACCUMULATE the parameter gradients into these fp32 [dim] buffers (both or neither)
    acc = dw_into is not None and db_into is not None
    dw = dw_into if acc else torch.empty(dim, dtype=torch.float32, device=x.device)
    db = db_into if acc else torch.empty(dim, dtype=torch.float32, device=x.device)
    ws = torch.empty(L.lib().mb_layernorm_bwd_workspace(rows, dim), dtype=torch.uint8, device=x.device)
    nbytes = rows * dim * (dy.element_size() + 4.0 + 4.0 + (4.0 if dres is not None else 0.0) +
                           (2.0 if want_bf16 else 0.0))
    with _rec("layernorm_bwd", nbytes, "byte", kernels=2):
        L.check(L.lib().mb_layernorm_bwd(dy.data_ptr(), L.MB_BF16 if dy.dtype == torch.bfloat16 else L.MB_F32,
                                         x.data_ptr(), weight.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                         _ptr(dres), dx.data_ptr(), _ptr(dxb), dw.data_ptr(), db.data_ptr(), 1 if acc else 0,
                                         ws.data_ptr(), rows, dim, x.stride(0), dy.stride(0), dx.stride(0),
                                         _ptr(dx_colsum), 1 if dx_colsum_accumulate else 0,
                                         _stream()), "mb_layernorm_bwd")
    if want_bf16:
        return dx, dxb, dw, db
    return dx, dw, db


def colsum(a: torch.Tensor, into: torch.Tensor | None = None) -> torch.Tensor:
    """fp32 column sums of a 2-D bf16/f32 matrix (bias gradient).  ``into``: accumulate into this
    fp32 [cols] buffer instead of returning a new one."""
    _req_cuda(a)
    rows, cols = a.shape
    out = into if into is not None else torch.empty(cols, dtype=torch.float32, device=a.device)
    ws = torch.empty(L.lib().mb_colsum_workspace(rows, cols), dtype=torch.uint8, device=a.device)
    with _rec("colsum", float(rows * cols * a.element_size()), "byte", kernels=2):
        L.check(L.lib().mb_colsum(a.data_ptr(), L.MB_BF16 if a.dtype == torch.bfloat16 else L.MB_F32,
                                  out.data_ptr(), 1 if into is not None else 0, ws.data_ptr(), rows, cols,
                                  a.stride(0), _stream()),
                "mb_colsum")
    return out


def token_gather(src: torch.Tensor, ids_keep: torch.Tensor, global_tokens: torch.Tensor) -> torch.Tensor:
    """[B, n_src, D] fp32, ids [B, n_keep] int64, global [n_glob, D] -> [B, n_keep + n_glob, D]."""
    _req_cuda(src, ids_keep, global_tokens)
    B, n_src, D = src.shape
    n_keep = ids_keep.shape[1]
    n_glob = global_tokens.shape[0]
    assert src.is_contiguous() and ids_keep.is_contiguous() and global_tokens.is_contiguous()
    assert src.dtype == torch.float32 and ids_keep.dtype == torch.int64
    out = torch.empty((B, n_keep + n_glob, D), dtype=torch.float32, device=src.device)
    with _rec("token_gather", 8.0 * B * (n_keep + n_glob) * D, "byte"):
        L.check(L.lib().mb_token_gather_fwd(src.data_ptr(), ids_keep.data_ptr(), global_tokens.data_ptr(),
                                            out.data_ptr(), B, n_src, n_keep, n_glob, D, _stream()),
                "mb_token_gather_fwd")
    return out


def token_gather_bwd(dout: torch.Tensor, ids_keep: torch.Tensor, n_src: int, n_glob: int):
    _req_cuda(dout, ids_keep)
    B, n_out, D = dout.shape
    n_keep = n_out - n_glob
    assert dout.is_contiguous() and dout.dtype == torch.float32
    dsrc = torch.empty((B, n_src, D), dtype=torch.float32, device=dout.device)
    dglob = torch.empty((n_glob, D), dtype=torch.float32, device=dout.device)
    with _rec("token_scatter", 4.0 * B * (n_src + 2 * n_keep) * D, "byte", kernels=3):
        L.check(L.lib().mb_token_gather_bwd(dout.data_ptr(), ids_keep.data_ptr(), dsrc.data_ptr(),
                                            dglob.data_ptr(), B, n_src, n_keep, n_glob, D, _stream()),
                "mb_token_gather_bwd")
    return dsrc, dglob


def fill_global_rows(global_tokens: torch.Tensor, out: torch.Tensor, row_offset: int):
    """out[:, row_offset:row_offset+n_glob, :] = global_tokens (out is [B, rows_total, D] fp32)."""
    _req_cuda(global_tokens, out)
    B, rows_total, D = out.shape
    with _rec("fill_global_rows", 4.0 * B * global_tokens.shape[0] * D, "byte"):
        L.check(L.lib().mb_fill_global_rows(global_tokens.data_ptr(), out.data_ptr(), B, rows_total,
                                            row_offset, global_tokens.shape[0], D, _stream()),
                "mb_fill_global_rows")
    return out


def cast_bf16(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """fp32 -> bf16; ``out``: an existing contiguous bf16 tensor of the same shape, rewritten in place."""
    _req_cuda(x, out)
    assert x.dtype == torch.float32 and x.is_contiguous()
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    else:
        assert out.dtype == torch.bfloat16 and out.is_contiguous() and out.numel() == x.numel()
    with _rec("cast_bf16", 6.0 * x.numel(), "byte"):
        L.check(L.lib().mb_cast_f32_to_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream()),
                "mb_cast_f32_to_bf16")
    return out


def attention_bwd(q, k, v, out, d_out, lse, dq, dk, dv, *, batch, heads, nq, nk, head_dim, scale):
    """Writes dq/dk/dv (bf16 views, token-major) given the forward tensors and d_out."""
    _req_cuda(q, k, v, out, d_out, lse, dq, dk, dv)
    a = L.AttnBwdArgs()
    a.q, a.k, a.v, a.out, a.d_out, a.lse = (q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(),
                                            d_out.data_ptr(), lse.data_ptr())
    a.dq, a.dk, a.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    wsb = L.lib().mb_attn_bwd_workspace(batch, heads, nq, nk, head_dim)
    ws = torch.empty(wsb, dtype=torch.uint8, device=q.device) if wsb else None
    a.workspace = _ptr(ws)
    a.batch, a.heads, a.nq, a.nk = batch, heads, nq, nk
    a.ldq, a.ldk, a.ldv, a.ldo, a.lddo = q.stride(0), k.stride(0), v.stride(0), out.stride(0), d_out.stride(0)
    a.lddq, a.lddk, a.lddv = dq.stride(0), dk.stride(0), dv.stride(0)
    a.head_dim = head_dim
    a.scale = scale
    with _rec("attn_bwd", 10.0 * batch * heads * nq * nk * head_dim, "flop"):
        L.check(L.lib().mb_attn_bwd(C.byref(a), _stream()), "mb_attn_bwd")


def semseg_patches(labels: torch.Tensor, class_emb_bf16: torch.Tensor, ph: int, pw: int,
                   row_src: torch.Tensor | None = None) -> torch.Tensor:
    """labels int64 [B,H,W] -> bf16 [B*(H/ph)*(W/pw), E*ph*pw] (embedding lookup + patch extraction).
    ``row_src`` (int32 [T]): only these source patches, in this order (negative: zero row) -> bf16 [T, E*ph*pw]."""
    _req_cuda(labels, class_emb_bf16, row_src)
    assert labels.dtype == torch.int64 and labels.is_contiguous()
    B, H, W = labels.shape
    n_cls, E = class_emb_bf16.shape
    rows = B * (H // ph) * (W // pw) if row_src is None else row_src.numel()
    out = torch.empty((rows, E * ph * pw), dtype=torch.bfloat16, device=labels.device)
    with _rec("semseg_patches", float(out.numel() * 2 + labels.numel() * 8), "byte"):
        L.check(L.lib().mb_semseg_patches_rows(labels.data_ptr(), class_emb_bf16.data_ptr(), out.data_ptr(),
                                               _ptr(row_src), rows, B, H, W, ph, pw, n_cls, E, _stream()),
                "mb_semseg_patches_rows")
    return out


def class_emb_grad(labels, d_patches_bf16, n_cls: int, E: int, ph: int, pw: int,
                   row_src: torch.Tensor | None = None) -> torch.Tensor:
    """``row_src`` (int32 [T]): gradient row r belongs to source patch row_src[r] (negative: skipped)."""
    _req_cuda(labels, d_patches_bf16, row_src)
    B, H, W = labels.shape
    out = torch.empty((n_cls, E), dtype=torch.float32, device=labels.device)
    rows = 0 if row_src is None else row_src.numel()
    with _rec("class_emb_grad", float(d_patches_bf16.numel() * 2 + labels.numel() * 8), "byte", kernels=2):
        L.check(L.lib().mb_class_emb_grad_rows(labels.data_ptr(), d_patches_bf16.data_ptr(), out.data_ptr(),
                                               _ptr(row_src), rows, B, H, W, ph, pw, n_cls, E, _stream()),
                "mb_class_emb_grad_rows")
    return out


# ---------------------------------------------------------------------------------------------
# visible-token embedding (csrc/visible.cu)
# ---------------------------------------------------------------------------------------------
def visible_rows(ids_keep: torch.Tensor, starts, counts, n_glob: int):
    """ids_keep int64 [B, n_keep] -> (row_src int32 [n_mod, T], row_cls int32 [T]), T = B * (n_keep + n_glob)."""
    _req_cuda(ids_keep)
    assert ids_keep.dtype == torch.int64 and ids_keep.is_contiguous()
    B, n_keep = ids_keep.shape
    n_mod = len(starts)
    T = B * (n_keep + n_glob)
    row_src = torch.empty((n_mod, T), dtype=torch.int32, device=ids_keep.device)
    row_cls = torch.empty((T,), dtype=torch.int32, device=ids_keep.device)
    st = (C.c_int32 * n_mod)(*[int(v) for v in starts])
    ct = (C.c_int32 * n_mod)(*[int(v) for v in counts])
    with _rec("visible_rows", float(ids_keep.numel() * 8 + (n_mod + 1) * T * 4), "byte"):
        L.check(L.lib().mb_visible_rows(ids_keep.data_ptr(), B, n_keep, n_glob, n_mod, st, ct, row_src.data_ptr(),
                                        row_cls.data_ptr(), _stream()), "mb_visible_rows")
    return row_src, row_cls


def embed_rows_init(row_src: torch.Tensor, row_cls: torch.Tensor, counts, biases, pos_rows,
                    global_tokens: torch.Tensor | None, dim: int) -> torch.Tensor:
    """fp32 [T, dim]: bias_m + pos_m[token] on the rows of modality m, global token g on the global rows."""
    _req_cuda(row_src, row_cls, global_tokens, *biases, *pos_rows)
    n_mod, T = row_src.shape
    out = torch.empty((T, dim), dtype=torch.float32, device=row_src.device)
    ct = (C.c_int32 * n_mod)(*[int(v) for v in counts])
    bp = (C.c_void_p * n_mod)(*[_ptr(b) for b in biases])
    pp = (C.c_void_p * n_mod)(*[_ptr(p) for p in pos_rows])
    for t in list(biases) + list(pos_rows) + [global_tokens]:
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous())
    with _rec("embed_rows_init", float(2 * T * dim * 4), "byte"):
        L.check(L.lib().mb_embed_rows_init(row_src.data_ptr(), row_cls.data_ptr(), n_mod, ct, bp, pp,
                                           _ptr(global_tokens), out.data_ptr(), T, dim, _stream()),
                "mb_embed_rows_init")
    return out


def gather_patches32(img: torch.Tensor, row_src: torch.Tensor, want_f32: bool = True, want_bf16: bool = False):
    """img fp32 [B, 1, H, W]; row_src int32 [T] -> (fp32 [T, 1024] | None, bf16 [T, 1024] | None)."""
    _req_cuda(img, row_src)
    assert img.dtype == torch.float32 and img.is_contiguous() and img.shape[1] == 1
    H, W = img.shape[-2:]
    T = row_src.numel()
    a32 = torch.empty((T, 1024), dtype=torch.float32, device=img.device) if want_f32 else None
    a16 = torch.empty((T, 1024), dtype=torch.bfloat16, device=img.device) if want_bf16 else None
    with _rec("gather_patches32", float(T * 1024 * (4 + (4 if want_f32 else 0) + (2 if want_bf16 else 0))), "byte"):
        L.check(L.lib().mb_gather_patches32(img.data_ptr(), row_src.data_ptr(), _ptr(a32), _ptr(a16), T, H, W,
                                            _stream()), "mb_gather_patches32")
    return a32, a16


def class_colsum(dy: torch.Tensor, row_cls: torch.Tensor, n_classes: int) -> torch.Tensor:
    """fp32 [n_classes, dim]: sums of the rows of dy (fp32 [T, dim]) by row class."""
    _req_cuda(dy, row_cls)
    assert dy.dtype == torch.float32 and dy.is_contiguous() and dy.dim() == 2
    T, dim = dy.shape
    out = torch.empty((n_classes, dim), dtype=torch.float32, device=dy.device)
    ws = torch.empty(L.lib().mb_class_colsum_workspace(T, dim, n_classes), dtype=torch.uint8, device=dy.device)
    with _rec("class_colsum", float(T * dim * 4), "byte", kernels=2):
        L.check(L.lib().mb_class_colsum(dy.data_ptr(), row_cls.data_ptr(), out.data_ptr(), ws.data_ptr(), T, dim,
                                        n_classes, _stream()), "mb_class_colsum")
    return out


def dec_assemble(ctx, mask_token, emb, ids_keep, ids_restore, q_start: int, n_q: int, n_glob: int):
    """ctx f32 [B, n_vis+n_glob, Dd] -> (queries f32 [B, n_q, Dd], context f32 [B, n_vis+n_glob, Dd])."""
    _req_cuda(ctx, mask_token, emb, ids_keep, ids_restore)
    B, n_ctx, Dd = ctx.shape
    n_vis = n_ctx - n_glob
    n_all = ids_restore.shape[1]
    q = torch.empty((B, n_q, Dd), dtype=torch.float32, device=ctx.device)
    c = torch.empty((B, n_ctx, Dd), dtype=torch.float32, device=ctx.device)
    with _rec("dec_assemble", 4.0 * Dd * B * (2 * n_q + 2 * n_ctx), "byte"):
        L.check(L.lib().mb_dec_assemble_fwd(ctx.data_ptr(), mask_token.data_ptr(), emb.data_ptr(),
                                            ids_keep.data_ptr(), ids_restore.data_ptr(), q.data_ptr(),
                                            c.data_ptr(), B, n_vis, n_glob, n_all, q_start, n_q, Dd,
                                            _stream()), "mb_dec_assemble_fwd")
    return q, c


def dec_assemble_bwd(dq, dc, ids_keep, ids_restore, q_start: int, n_glob: int):
    _req_cuda(dq, dc, ids_keep, ids_restore)
    B, n_q, Dd = dq.shape
    n_ctx = dc.shape[1]
    n_vis = n_ctx - n_glob
    n_all = ids_restore.shape[1]
    dctx = torch.empty((B, n_ctx, Dd), dtype=torch.float32, device=dq.device)
    demb = torch.empty((n_all, Dd), dtype=torch.float32, device=dq.device)
    dmask = torch.empty((Dd,), dtype=torch.float32, device=dq.device)
    ws = torch.empty((n_q, Dd), dtype=torch.float32, device=dq.device)
    with _rec("dec_assemble_bwd", 4.0 * Dd * B * (2 * n_q + 3 * n_ctx), "byte", kernels=3):
        L.check(L.lib().mb_dec_assemble_bwd(dq.data_ptr(), dc.data_ptr(), ids_keep.data_ptr(),
                                            ids_restore.data_ptr(), dctx.data_ptr(), demb.data_ptr(),
                                            dmask.data_ptr(), ws.data_ptr(), B, n_vis, n_glob, n_all,
                                            q_start, n_q, Dd, _stream()), "mb_dec_assemble_bwd")
    return dctx, demb, dmask


def patchify_cast(img: torch.Tensor, ph: int, pw: int) -> torch.Tensor:
    """f32 [B,C,H,W] -> bf16 [B*(H/ph)*(W/pw), C*ph*pw]."""
    _req_cuda(img)
    assert img.dtype == torch.float32 and img.is_contiguous()
    B, Cc, H, W = img.shape
    gh, gw = H // ph, W // pw
    out = torch.empty((B * gh * gw, Cc * ph * pw), dtype=torch.bfloat16, device=img.device)
    with _rec("patchify_cast", 6.0 * img.numel(), "byte"):
        L.check(L.lib().mb_patchify_cast(img.data_ptr(), out.data_ptr(), B, Cc, ph, pw, gh, gw, _stream()),
                "mb_patchify_cast")
    return out


def masked_mse_fwd(pred, target, mask, scale: int):
    _req_cuda(pred, target, mask)
    B, Cc, H, W = pred.shape
    loss = torch.empty((), dtype=torch.float32, device=pred.device)
    coef = torch.empty((B,), dtype=torch.float32, device=pred.device)
    ws = torch.empty(L.lib().mb_masked_loss_workspace(B, H, W), dtype=torch.uint8, device=pred.device)
    with _rec("masked_mse_fwd", 8.0 * pred.numel(), "byte", kernels=2):
        L.check(L.lib().mb_masked_mse_fwd(pred.data_ptr(), target.data_ptr(), _ptr(mask), loss.data_ptr(),
                                          coef.data_ptr(), ws.data_ptr(), B, Cc, H, W, scale, _stream()),
                "mb_masked_mse_fwd")
    return loss, coef


def masked_mse_bwd(pred, target, mask, coef, gout, scale: int):
    B, Cc, H, W = pred.shape
    dpred = torch.empty_like(pred)
    with _rec("masked_mse_bwd", 12.0 * pred.numel(), "byte"):
        L.check(L.lib().mb_masked_mse_bwd(pred.data_ptr(), target.data_ptr(), _ptr(mask), coef.data_ptr(),
                                          gout.data_ptr(), dpred.data_ptr(), B, Cc, H, W, scale, _stream()),
                "mb_masked_mse_bwd")
    return dpred


def masked_ce_fwd(logits, target, mask, scale: int, smoothing: float):
    _req_cuda(logits, target, mask)
    B, Cc, H, W = logits.shape
    loss = torch.empty((), dtype=torch.float32, device=logits.device)
    coef = torch.empty((B,), dtype=torch.float32, device=logits.device)
    ws = torch.empty(L.lib().mb_masked_loss_workspace(B, H, W), dtype=torch.uint8, device=logits.device)
    with _rec("masked_ce_fwd", 4.0 * logits.numel() + 8.0 * target.numel(), "byte", kernels=2):
        L.check(L.lib().mb_masked_ce_fwd(logits.data_ptr(), target.data_ptr(), _ptr(mask), loss.data_ptr(),
                                         coef.data_ptr(), ws.data_ptr(), B, Cc, H, W, scale, smoothing,
                                         _stream()), "mb_masked_ce_fwd")
    return loss, coef


def masked_ce_bwd(logits, target, mask, coef, gout, scale: int, smoothing: float):
    B, Cc, H, W = logits.shape
    dl = torch.empty_like(logits)
    with _rec("masked_ce_bwd", 8.0 * logits.numel() + 8.0 * target.numel(), "byte"):
        L.check(L.lib().mb_masked_ce_bwd(logits.data_ptr(), target.data_ptr(), _ptr(mask), coef.data_ptr(),
                                         gout.data_ptr(), dl.data_ptr(), B, Cc, H, W, scale, smoothing,
                                         _stream()), "mb_masked_ce_bwd")
    return dl


# ---------------------------------------------------------------------------------------------
# LayerNorm + token mean-pool (classification tail)
# ---------------------------------------------------------------------------------------------
def ln_meanpool_fwd(x: torch.Tensor, gamma, beta, eps: float, row_begin: int, row_end: int,
                    pooled: torch.Tensor | None = None, col_offset: int = 0):
    """x f32 [B, N, D] -> pooled f32 [B, D] = mean over rows [row_begin, row_end) of LayerNorm(x).
    ``pooled`` / ``col_offset``: write into columns [col_offset, col_offset + D) of a wider [B, k*D] buffer.
    Returns (pooled, xhat_mean, mean, rstd)."""
    _req_cuda(x, gamma, beta, pooled)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 3
    B, N, D = x.shape
    if pooled is None:
        pooled = torch.empty((B, D), dtype=torch.float32, device=x.device)
    assert pooled.dtype == torch.float32 and pooled.stride(1) == 1
    xhat_mean = torch.empty((B, D), dtype=torch.float32, device=x.device)
    mean = torch.empty((B, N), dtype=torch.float32, device=x.device)
    rstd = torch.empty((B, N), dtype=torch.float32, device=x.device)
    ws = torch.empty(L.lib().mb_ln_meanpool_workspace(B, D), dtype=torch.uint8, device=x.device)
    with _rec("ln_meanpool_fwd", 4.0 * B * (row_end - row_begin) * D, "byte", kernels=2):
        L.check(L.lib().mb_ln_meanpool_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                           pooled.data_ptr() + 4 * col_offset, pooled.stride(0),
                                           xhat_mean.data_ptr(), mean.data_ptr(), rstd.data_ptr(), ws.data_ptr(),
                                           B, N, D, row_begin, row_end, eps, _stream()), "mb_ln_meanpool_fwd")
    return pooled, xhat_mean, mean, rstd


def ln_meanpool_bwd(d_pooled: torch.Tensor, col_offset: int, x, gamma, mean, rstd, xhat_mean, row_begin: int,
                    row_end: int, dx: torch.Tensor | None = None, d_gamma=None, d_beta=None,
                    accumulate: bool = False):
    """Returns (dx f32 [B, N, D], d_gamma, d_beta).  With ``dx`` given, only the pooled rows are written
    (second range of the token_mix pooling); ``accumulate`` adds into d_gamma / d_beta."""
    _req_cuda(d_pooled, x, gamma, mean, rstd, xhat_mean, dx)
    B, N, D = x.shape
    assert d_pooled.dtype == torch.float32 and d_pooled.stride(1) == 1
    zero_outside = dx is None
    if dx is None:
        dx = torch.empty_like(x)
    if d_gamma is None:
        assert not accumulate
        d_gamma = torch.empty(D, dtype=torch.float32, device=x.device)
        d_beta = torch.empty(D, dtype=torch.float32, device=x.device)
    with _rec("ln_meanpool_bwd", 8.0 * B * (row_end - row_begin) * D + (4.0 * B * N * D if zero_outside else 0.0),
              "byte", kernels=2):
        L.check(L.lib().mb_ln_meanpool_bwd(d_pooled.data_ptr() + 4 * col_offset, d_pooled.stride(0), x.data_ptr(),
                                           gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(), xhat_mean.data_ptr(),
                                           dx.data_ptr(), d_gamma.data_ptr(), d_beta.data_ptr(),
                                           1 if accumulate else 0, 1 if zero_outside else 0, B, N, D, row_begin,
                                           row_end, _stream()), "mb_ln_meanpool_bwd")
    return dx, d_gamma, d_beta


# ---------------------------------------------------------------------------------------------
# on-device mask sampling
# ---------------------------------------------------------------------------------------------
def sample_masks(seed: int, draw_counter: torch.Tensor, done_counter: torch.Tensor, counts, alphas,
                 batch: int, n_encoded: int, uniform_tasks: bool = False):
    """Returns (mask_all i64 [B, n_all], ids_keep i64 [B, n_encoded], ids_restore i64 [B, n_all]).
    ``draw_counter``: 1-element int64 CUDA tensor (advanced by the kernel); ``done_counter``: 1-element
    int32 CUDA tensor, zero-initialised scratch."""
    _req_cuda(draw_counter, done_counter)
    assert draw_counter.dtype == torch.int64 and done_counter.dtype == torch.int32
    dev = draw_counter.device
    n_all = int(sum(counts))
    mask_all = torch.empty((batch, n_all), dtype=torch.int64, device=dev)
    keep = torch.empty((batch, n_encoded), dtype=torch.int64, device=dev)
    restore = torch.empty((batch, n_all), dtype=torch.int64, device=dev)
    c_counts = (C.c_int32 * len(counts))(*[int(c) for c in counts])
    c_alphas = (C.c_float * len(counts))(*[float(a) for a in alphas])
    with _rec("sample_masks", 8.0 * batch * (2 * n_all + n_encoded), "byte"):
        L.check(L.lib().mb_sample_masks(int(seed) & 0xFFFFFFFFFFFFFFFF, draw_counter.data_ptr(),
                                        done_counter.data_ptr(), c_counts, c_alphas, len(counts), batch, n_encoded,
                                        1 if uniform_tasks else 0, mask_all.data_ptr(), keep.data_ptr(),
                                        restore.data_ptr(), _stream()), "mb_sample_masks")
    return mask_all, keep, restore


# ---------------------------------------------------------------------------------------------
# device-side input pipeline
# ---------------------------------------------------------------------------------------------
def augment_image(src_u8: torch.Tensor, params: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """uint8 [B, H, W] + params f32 [B, 8] -> f32 [B, 1, H, W] (flip, intensity shift, affine warp)."""
    _req_cuda(src_u8, params, out)
    assert src_u8.dtype == torch.uint8 and src_u8.is_contiguous() and src_u8.dim() == 3
    assert params.dtype == torch.float32 and params.is_contiguous() and params.shape == (src_u8.shape[0], 8)
    B, H, W = src_u8.shape
    if out is None:
        out = torch.empty((B, 1, H, W), dtype=torch.float32, device=src_u8.device)
    with _rec("augment_image", 5.0 * B * H * W, "byte"):
        L.check(L.lib().mb_augment_image(src_u8.data_ptr(), params.data_ptr(), out.data_ptr(), B, H, W, _stream()),
                "mb_augment_image")
    return out


def augment_labels(src_u8: torch.Tensor, params: torch.Tensor, out_hw, out: torch.Tensor | None = None):
    """uint8 class map [B, H, W] + params f32 [B, 8] -> int64 [B, OH, OW]."""
    _req_cuda(src_u8, params, out)
    assert src_u8.dtype == torch.uint8 and src_u8.is_contiguous() and src_u8.dim() == 3
    assert params.dtype == torch.float32 and params.is_contiguous() and params.shape == (src_u8.shape[0], 8)
    B, H, W = src_u8.shape
    OH, OW = out_hw
    if out is None:
        out = torch.empty((B, OH, OW), dtype=torch.int64, device=src_u8.device)
    with _rec("augment_labels", 1.0 * B * H * W / max(1, (H // OH) * (W // OW)) * 4 + 8.0 * B * OH * OW, "byte"):
        L.check(L.lib().mb_augment_labels(src_u8.data_ptr(), params.data_ptr(), out.data_ptr(), B, H, W, OH, OW,
                                          _stream()), "mb_augment_labels")
    return out
