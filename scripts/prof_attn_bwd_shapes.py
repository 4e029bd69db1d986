"""attn_bwd at the cfg-4 (N = 99) and cfg-5 (N = 257) shapes, three launches each (target of an ncu capture)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mirage_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
for B, H, N in ((256, 16, 99), (64, 16, 257)):
    D = H * 64
    qkv = torch.randn(B * N, 3 * D, device=dev).bfloat16()
    lse = torch.empty(B, H, N, device=dev)
    o = ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], batch=B, heads=H, nq=N, nk=N, head_dim=64, scale=0.125, lse=lse)
    do = torch.randn_like(o)
    dqkv = torch.empty_like(qkv)
    for _ in range(3):
        ops.attention_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, do, lse, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:],
                          batch=B, heads=H, nq=N, nk=N, head_dim=64, scale=0.125)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.attention_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, do, lse, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:],
                          batch=B, heads=H, nq=N, nk=N, head_dim=64, scale=0.125)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 200
    print(f"attn_bwd B={B} H={H} N={N}: {us:.1f} us  {10.0 * B * H * N * N * 64 / us / 1e6:.0f} TFLOP/s", flush=True)
