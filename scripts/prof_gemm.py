"""One GEMM shape for ncu: python scripts/prof_gemm.py M N K [residual]"""
import sys
import torch
sys.path.insert(0, ".")
from mirage_b200 import ops
m, n, k = map(int, sys.argv[1:4])
res = len(sys.argv) > 4
dev = torch.device("cuda:0")
a = torch.randn(m, k, device=dev).bfloat16()
w = (torch.randn(n, k, device=dev) * k ** -0.5).bfloat16()
bias = torch.randn(n, device=dev)
r = torch.randn(m, n, device=dev) if res else None
out = torch.empty(m, n, dtype=torch.float32 if res else torch.bfloat16, device=dev)
for _ in range(3):
    ops.gemm(a, w, m=m, n=n, k=k, bias=bias, residual=r, out=out)
torch.cuda.synchronize()
