"""Small kernels of the cfg-4 step at their real shapes, a few launches each (target of an ncu capture):
    ncu --set full --clock-control none -k regex:'layernorm_bwd|attn_fwd_kernel|colsum_partial|layernorm_fwd' \
        -c 8 -o gpurun_out/r02_rowkernels python scripts/prof_rowkernels.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mirage_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
M, D, B, H, N = 256 * 99, 1024, 256, 16, 99
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, D, device=dev, generator=g)
w = torch.ones(D, device=dev)
b = torch.zeros(D, device=dev)
dy = torch.randn(M, D, device=dev, generator=g).bfloat16()
dres = torch.randn(M, D, device=dev, generator=g)
cs = torch.zeros(D, device=dev)
qkv = torch.randn(M, 3 * D, device=dev, generator=g).bfloat16()


def timed(name, fn, nbytes, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{name:28s} {us:8.1f} us  {nbytes / us / 1e6:7.2f} TB/s", flush=True)


y, mean, rstd = ops.layernorm(x, w, b, 1e-6, save_stats=True)
timed("layernorm_fwd", lambda: ops.layernorm(x, w, b, 1e-6, save_stats=True), M * D * 6)
timed("layernorm_bwd (+dres,bf16,cs)", lambda: ops.layernorm_bwd(dy, x, w, mean, rstd, dres=dres, want_bf16=True,
                                                                dx_colsum=cs), M * D * 16)
lse = torch.empty(B, H, N, device=dev)
timed("attn_fwd N=99", lambda: ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], batch=B, heads=H, nq=N, nk=N,
                                             head_dim=64, scale=0.125, lse=lse), M * D * 2 * 4)
timed("colsum bf16 [M,3D]", lambda: ops.colsum(qkv), M * 3 * D * 2)
