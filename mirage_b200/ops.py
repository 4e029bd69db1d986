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
         img_hw: tuple[int, int] | None = None) -> torch.Tensor:
    """out[m, n] = epilogue(A * B^T); see ``mb_gemm`` in include/mirage_b200.h for the contract.

    ``a`` / ``b`` are 2-D (or, for MB_A_PATCH32, the [B,1,H,W] fp32 image batch) with unit stride in
    the last dimension; leading dimensions are taken from ``stride(0)``.
    """
    _req_cuda(a, b, out, bias, residual, aux_out, dgelu_aux)
    tf32 = a.dtype == torch.float32
    if a.dtype not in (torch.bfloat16, torch.float32) or b.dtype != a.dtype:
        raise L.MirageB200Error(f"gemm: unsupported operand dtypes {a.dtype}, {b.dtype}")
    if out is None:
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
    args.ldc = out.stride(0)
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
    L.check(L.lib().mb_gemm(C.byref(args), _stream()), "mb_gemm")
    return out
