"""GPU check: HF-style MIRAGEWrapper (MIRAGELight encoder) vs the CPU oracle, plus a first timing."""
import sys
import time

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from helpers import load_synth, parity, synth_images  # noqa: E402
from mirage_b200.mirage_hf import MIRAGEWrapper  # noqa: E402
from oracle import mirage_oracle as O  # noqa: E402

dev = torch.device("cuda:0")


def check(size, B):
    m = MIRAGEWrapper(size=size)
    sd = load_synth(m.model, seed=0)
    m = m.to(dev).eval()
    x = synth_images(B, ["bscan", "slo"], seed=1234)
    depth, heads = (12, 12) if size == "base" else (24, 16)
    t0 = time.time()
    with torch.no_grad():
        ref = O.light_forward(x, sd, depth, heads)
    t_cpu = time.time() - t0
    with torch.no_grad():
        out = m({k: v.to(dev) for k, v in x.items()})
    torch.cuda.synchronize()
    pm = parity(out, ref)
    ok = pm["max_rel"] <= 2e-2 and pm["min_cos"] >= 0.999
    print(f"[{'PASS' if ok else 'FAIL'}] encoder {size} B={B}: out {tuple(out.shape)} {pm}  (oracle cpu {t_cpu:.2f}s)",
          flush=True)
    # layer-by-layer drift, useful when the end-to-end check fails
    if not ok:
        with torch.no_grad():
            refs = O.light_forward(x, sd, depth, heads, return_all_layers=True)
            outs = m.model({k: v.to(dev) for k, v in x.items()}, return_all_layers=True)
        for i, (o, r) in enumerate(zip(outs, refs)):
            print(f"    layer {i}: {parity(o, r)}")
    return ok


def perf(size, B, iters=5):
    m = MIRAGEWrapper(size=size).to(dev).eval()
    x = {k: v.to(dev) for k, v in synth_images(8, ["bscan", "slo"]).items()}
    x = {k: v.repeat(B // 8, 1, 1, 1).contiguous() for k, v in x.items()}
    with torch.no_grad():
        for _ in range(2):
            m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            m(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    gf = {"base": 97.65, "large": 336.79}[size]
    print(f"[PERF] encoder {size} B={B}: {ms:.2f} ms/step = {B / ms * 1e3:.0f} img/s = "
          f"{B * gf / ms:.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    ok = check("base", 1)
    ok &= check("base", 3)
    ok &= check("large", 2)
    perf("base", 256)
    perf("large", 256)
    sys.exit(0 if ok else 1)
