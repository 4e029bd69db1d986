"""GPU check of the folded-LayerNorm GEMM epilogues (EPI_RES_LN producer, EPI_BF16_LN / EPI_GELU_LN consumers)."""
import sys
import torch
sys.path.insert(0, ".")
from mirage_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
ok = True
for (T, D, N2) in [(1026, 768, 2304), (4104, 1024, 4096), (300, 256, 1024)]:
    a = torch.randn(T, D, device=dev).bfloat16()
    w = (torch.randn(D, D, device=dev) * D ** -0.5).bfloat16()
    b = torch.randn(D, device=dev) * 0.1
    res = torch.randn(T, D, device=dev) * 2 + 0.3
    stats = torch.empty(T, 1 + ops.gemm_ln_parts(T, D), 2, device=dev)
    twin = torch.empty(T, D, dtype=torch.bfloat16, device=dev)
    out = ops.gemm(a, w, m=T, n=D, k=D, bias=b, residual=res, out_dtype=torch.float32, twin_out=twin, row_stats=stats)
    ref = a.float() @ w.float().t() + b + res
    e_out = (out - ref).abs().max().item()
    e_twin = (twin.float() - out.to(torch.bfloat16).float()).abs().max().item()
    e_s1 = ((stats[:, 0, 0] - ref.sum(1)).abs().max() / ref.sum(1).abs().max()).item()
    e_s2 = ((stats[:, 0, 1] - (ref * ref).sum(1)).abs().max() / (ref * ref).sum(1).abs().max()).item()
    print(f"producer T={T} D={D}: out {e_out:.2e} twin {e_twin:.2e} sum {e_s1:.2e} sumsq {e_s2:.2e}")
    ok &= e_out < 2e-2 and e_twin == 0.0 and e_s1 < 1e-4 and e_s2 < 1e-4
    # consumer
    gamma = 1 + 0.1 * torch.randn(D, device=dev)
    beta = 0.1 * torch.randn(D, device=dev)
    w2 = torch.randn(N2, D, device=dev) * D ** -0.5
    b2 = torch.randn(N2, device=dev) * 0.1
    wp = (w2 * gamma[None, :]).to(torch.bfloat16).contiguous()
    c1 = wp.float().sum(1).contiguous()
    c2 = (b2 + w2 @ beta).contiguous()
    want = torch.nn.functional.layer_norm(out, (D,), gamma, beta, 1e-6) @ w2.t() + b2
    for gelu in (False, True):
        y = ops.gemm(twin, wp, m=T, n=N2, k=D, bias=c2, gelu=gelu, ln_stats=stats, ln_c1=c1, ln_eps=1e-6)
        r = torch.nn.functional.gelu(want) if gelu else want
        err = ((y.float() - r).abs().max() / r.abs().max()).item()
        print(f"  consumer gelu={gelu} N={N2}: max_rel {err:.3e}")
        ok &= err < 2e-2
print("ALL PASS" if ok else "FAIL")
