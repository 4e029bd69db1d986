"""Per-shape cost of the folded-LayerNorm epilogues vs the plain ones (cfg 2 shapes, CUDA events)."""
import sys
import torch
sys.path.insert(0, ".")
from mirage_b200 import ops
dev = torch.device("cuda:0")
T, D = 131328, 1024


def t(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


a = torch.randn(T, D, device=dev).bfloat16()
a4 = torch.randn(T, 4 * D, device=dev).bfloat16()
res = torch.randn(T, D, device=dev)
stats = torch.empty(T, 1 + ops.gemm_ln_parts(T, D), 2, device=dev)
twin = torch.empty(T, D, dtype=torch.bfloat16, device=dev)
b1 = torch.randn(D, device=dev)
out32 = torch.empty(T, D, device=dev)
for (name, A, K) in (("proj", a, D), ("fc2", a4, 4 * D)):
    w = torch.randn(D, K, device=dev).bfloat16()
    plain = t(lambda: ops.gemm(A, w, m=T, n=D, k=K, bias=b1, residual=res, out=out32))
    fold = t(lambda: ops.gemm(A, w, m=T, n=D, k=K, bias=b1, residual=res, out=out32, twin_out=twin, row_stats=stats))
    print(f"[PERF] {name}: EPI_RES {plain:.0f} us | EPI_RES_LN {fold:.0f} us")
for (name, N, gelu) in (("qkv", 3 * D, False), ("fc1", 4 * D, True)):
    w = torch.randn(N, D, device=dev).bfloat16()
    bn = torch.randn(N, device=dev)
    c1 = torch.randn(N, device=dev)
    o = torch.empty(T, N, dtype=torch.bfloat16, device=dev)
    plain = t(lambda: ops.gemm(a, w, m=T, n=N, k=D, bias=bn, gelu=gelu, out=o))
    fold = t(lambda: ops.gemm(a, w, m=T, n=N, k=D, bias=bn, gelu=gelu, out=o, ln_stats=stats, ln_c1=c1))
    print(f"[PERF] {name}: plain {plain:.0f} us | LN-folded {fold:.0f} us")
x = torch.randn(T, D, device=dev)
g = torch.ones(D, device=dev)
print(f"[PERF] standalone layernorm: {t(lambda: ops.layernorm(x, g, b1)):.0f} us")
