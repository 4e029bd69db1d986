"""Hunt for intermittent non-finite outputs: encoder large B=2, per-block finiteness, repeated."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import load_synth, synth_images
from mirage_b200 import ops
from mirage_b200.mirage_hf import MIRAGEWrapper

dev = torch.device("cuda:0")
# poison fresh memory so that any read of uninitialised data shows
junk = torch.full((1 << 28,), float("nan"), device=dev)
del junk
m = MIRAGEWrapper(size="large")
load_synth(m.model, seed=0)
m = m.to(dev).eval()
bad = []
def hook(i):
    def f(mod, inp, out):
        if not torch.isfinite(out).all():
            bad.append(i)
    return f
for i, blk in enumerate(m.model.encoder):
    blk.register_forward_hook(hook(i))
for B in (2, 1, 3, 2, 5, 2):
    x = {k: v.to(dev) for k, v in synth_images(B, ["bscan", "slo"], seed=1234).items()}
    for it in range(6):
        bad.clear()
        junk = torch.full((1 << 27,), float("nan"), device=dev); del junk
        with torch.no_grad():
            out = m(x)
        torch.cuda.synchronize()
        print(f"B={B} it={it} finite={bool(torch.isfinite(out).all())} first_bad_block={bad[:1]}", flush=True)
# attention alone
for B, H, n in [(2, 16, 513), (1, 16, 513), (3, 12, 513), (2, 16, 257)]:
    D = H * 64
    nbad = 0
    for it in range(20):
        junk = torch.full((1 << 26,), float("nan"), device=dev); del junk
        qkv = torch.randn(B * n, 3 * D, device=dev).bfloat16()
        o = ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], batch=B, heads=H, nq=n, nk=n, head_dim=64, scale=0.125)
        torch.cuda.synchronize()
        if not torch.isfinite(o.float()).all():
            nbad += 1
            rows = (~torch.isfinite(o.float())).any(1).nonzero().flatten()
            print("  bad rows", rows[:10].tolist(), "n", rows.numel())
    print(f"attn B={B} H={H} n={n}: bad {nbad}/20", flush=True)
