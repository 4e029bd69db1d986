import sys, torch
sys.path.insert(0, ".")
from mirage_b200 import ops
dev = torch.device("cuda:0")
B, H, n, hd = int(sys.argv[1]), 16, 513, 64
D = H * hd
qkv = torch.randn(B * n, 3 * D, device=dev).bfloat16()
o = ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], batch=B, heads=H, nq=n, nk=n, head_dim=hd, scale=hd ** -0.5)
torch.cuda.synchronize()
q = qkv[:, :D].float().reshape(B, n, H, hd).transpose(1, 2)
k = qkv[:, D:2 * D].float().reshape(B, n, H, hd).transpose(1, 2)
v = qkv[:, 2 * D:].float().reshape(B, n, H, hd).transpose(1, 2)
ref = (torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5, -1) @ v).transpose(1, 2).reshape(B * n, D)
print("B", B, "max err", (o.float() - ref).abs().max().item(), flush=True)
