"""Representative GEMMs of one ViT-L block at cfg-4 size (256 x 99 token rows), for ncu:
   ncu --set full -k regex:gemm_pair -s 12 -c 4 ... python scripts/prof_pretrain_gemms.py"""
import sys
import torch
sys.path.insert(0, ".")
from mirage_b200 import _lib as L
from mirage_b200 import ops
dev = torch.device("cuda:0")
t, d = 256 * 99, 1024
x = torch.randn(t, d, device=dev).bfloat16()
h = torch.randn(t, 4 * d, device=dev).bfloat16()
dy = torch.randn(t, d, device=dev).bfloat16()
res = torch.randn(t, d, device=dev)
w_qkv = (torch.randn(3 * d, d, device=dev) * d ** -0.5).bfloat16()
w_fc2 = (torch.randn(d, 4 * d, device=dev) * (4 * d) ** -0.5).bfloat16()
b = torch.randn(4 * d, device=dev)
dw = torch.zeros(4 * d, d, device=dev)
cs = torch.zeros(4 * d, device=dev)
def run():
    ops.gemm(x, w_qkv, m=t, n=3 * d, k=d, bias=b[:3 * d])                                        # fwd qkv
    ops.gemm(h, w_fc2, m=t, n=d, k=4 * d, bias=b[:d], residual=res, out=res)                      # fwd fc2 + residual
    ops.gemm(dy, w_fc2, m=t, n=4 * d, k=d, b_layout=L.MB_MAJOR_MN, dgelu_aux=h, colsum_out=cs)    # dgrad through GELU
    ops.gemm(h, x, m=4 * d, n=d, k=t, a_layout=L.MB_MAJOR_MN, b_layout=L.MB_MAJOR_MN, out=dw, atomic=True)  # wgrad fc1
for _ in range(4):
    run()
torch.cuda.synchronize()
