"""The four GEMMs of one ViT-L encoder block at cfg-2 size (256 x 513 tokens), for ncu:
   ncu --set full -k regex:gemm_pair -s 12 -c 4 ... python scripts/prof_encoder_gemms.py"""
import sys
import torch
sys.path.insert(0, ".")
from mirage_b200 import ops
dev = torch.device("cuda:0")
m, d = 256 * 513, 1024
x = torch.randn(m, d, device=dev).bfloat16()
h = torch.randn(m, 4 * d, device=dev).bfloat16()
res = torch.randn(m, d, device=dev)
w = {n: (torch.randn(o, i, device=dev) * i ** -0.5).bfloat16() for n, o, i in
     [("qkv", 3 * d, d), ("proj", d, d), ("fc1", 4 * d, d), ("fc2", d, 4 * d)]}
b = {n: torch.randn(v.shape[0], device=dev) for n, v in w.items()}
def run():
    ops.gemm(x, w["qkv"], m=m, n=3 * d, k=d, bias=b["qkv"])
    ops.gemm(x, w["proj"], m=m, n=d, k=d, bias=b["proj"], residual=res, out=res)
    ops.gemm(x, w["fc1"], m=m, n=4 * d, k=d, bias=b["fc1"], gelu=True)
    ops.gemm(h, w["fc2"], m=m, n=d, k=4 * d, bias=b["fc2"], residual=res, out=res)
for _ in range(4):
    run()
torch.cuda.synchronize()
