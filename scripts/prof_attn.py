import sys, torch
sys.path.insert(0, ".")
from mirage_b200 import ops
dev = torch.device("cuda:0")
B, H, n, hd = 256, 16, 513, 64
D = H * hd
qkv = torch.randn(B * n, 3 * D, device=dev).bfloat16()
out = torch.empty(B * n, D, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], batch=B, heads=H, nq=n, nk=n, head_dim=hd, scale=hd ** -0.5, out=out)
torch.cuda.synchronize()
