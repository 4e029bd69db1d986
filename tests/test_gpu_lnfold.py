"""-m gpu: LayerNorm folded into the GEMMs around it (functional.LN_FOLD, opt-in inference path): the producer
epilogue's bf16 twin / row statistics and the consumer epilogues against torch, and the ViT-L encoder with the fold
enabled against the fp32 oracle (same tolerance as the default path)."""
import pytest
import torch

from helpers import assert_parity, load_synth, synth_images

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,D,N2", [(1026, 768, 2304), (513, 1024, 4096), (300, 256, 1024)])
def test_lnfold_epilogues(T, D, N2):
    from mirage_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(T)
    a = torch.randn(T, D, device=dev, generator=g).bfloat16()
    w = (torch.randn(D, D, device=dev, generator=g) * D ** -0.5).bfloat16()
    b = torch.randn(D, device=dev, generator=g) * 0.1
    res = torch.randn(T, D, device=dev, generator=g) * 2 + 0.3
    parts = ops.gemm_ln_parts(T, D)
    stats = torch.full((2, T, 1 + parts, 2), float("nan"), device=dev)[1]   # every slot must be written (no zero-init)
    twin = torch.empty(T, D, dtype=torch.bfloat16, device=dev)
    out = ops.gemm(a, w, m=T, n=D, k=D, bias=b, residual=res, out_dtype=torch.float32, twin_out=twin, row_stats=stats)
    ref = a.float() @ w.float().t() + b + res
    assert (out - ref).abs().max().item() <= 1e-3
    assert torch.equal(twin, out.to(torch.bfloat16))
    tot = stats[:, 0]
    assert torch.allclose(tot, stats[:, 1:].sum(1), rtol=1e-5, atol=1e-3)
    assert ((tot[:, 0] - out.sum(1)).abs().max() / out.sum(1).abs().max()).item() <= 1e-5
    assert ((tot[:, 1] - (out * out).sum(1)).abs().max() / (out * out).sum(1).abs().max()).item() <= 1e-5
    # deterministic: a second launch gives the same bits (the first version accumulated with atomics)
    stats2 = torch.empty_like(stats)
    ops.gemm(a, w, m=T, n=D, k=D, bias=b, residual=res, out_dtype=torch.float32, twin_out=torch.empty_like(twin),
             row_stats=stats2)
    assert torch.equal(stats, stats2)
    gamma = 1 + 0.1 * torch.randn(D, device=dev, generator=g)
    beta = 0.1 * torch.randn(D, device=dev, generator=g)
    w2 = torch.randn(N2, D, device=dev, generator=g) * D ** -0.5
    b2 = torch.randn(N2, device=dev, generator=g) * 0.1
    wp = (w2 * gamma[None, :]).to(torch.bfloat16).contiguous()
    c1, c2 = wp.float().sum(1).contiguous(), (b2 + w2 @ beta).contiguous()
    want = torch.nn.functional.layer_norm(out, (D,), gamma, beta, 1e-6) @ w2.t() + b2
    for gelu in (False, True):
        y = ops.gemm(twin, wp, m=T, n=N2, k=D, bias=c2, gelu=gelu, ln_stats=stats, ln_c1=c1, ln_eps=1e-6)
        r = torch.nn.functional.gelu(want) if gelu else want
        assert ((y.float() - r).abs().max() / r.abs().max()).item() <= 2e-2


def test_encoder_with_lnfold_vs_oracle():
    from mirage_b200 import functional as Fn
    from mirage_b200.mirage_hf import MIRAGEWrapper
    from oracle import mirage_oracle as O
    dev = torch.device("cuda:0")
    m = MIRAGEWrapper(size="large")
    sd = load_synth(m.model, seed=0)
    m = m.to(dev).eval()
    x = synth_images(2, ["bscan", "slo"], seed=1234)
    prev = Fn.LN_FOLD
    Fn.LN_FOLD = True
    try:
        with torch.no_grad():
            out = m({k: v.to(dev) for k, v in x.items()})
    finally:
        Fn.LN_FOLD = prev
    with torch.no_grad():
        ref = O.light_forward(x, sd, 24, 16)
    assert_parity(out, ref, "ViT-L encoder with folded LayerNorm")
