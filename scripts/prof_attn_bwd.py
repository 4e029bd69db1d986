import sys, torch
sys.path.insert(0, ".")
from mirage_b200 import ops
dev = torch.device("cuda:0")
B, H, n, hd = 256, 16, 99, 64
D = H * hd
qkv = torch.randn(B * n, 3 * D, device=dev).bfloat16()
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
lse = torch.empty(B, H, n, device=dev)
o = ops.attention(q, k, v, batch=B, heads=H, nq=n, nk=n, head_dim=hd, scale=hd ** -0.5, lse=lse)
do = torch.randn(B * n, D, device=dev).bfloat16()
dqkv = torch.empty_like(qkv)
for _ in range(3):
    ops.attention_bwd(q, k, v, o, do, lse, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], batch=B, heads=H, nq=n, nk=n, head_dim=hd, scale=hd ** -0.5)
torch.cuda.synchronize()
