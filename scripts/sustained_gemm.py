"""Sustained (power-capped) GEMM throughput: mb_gemm vs torch.matmul (cuBLAS) in back-to-back loops of a few
seconds each, with SM clock / power sampled by nvidia-smi during the loop."""
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
from mirage_b200 import _lib as L
from mirage_b200 import ops

dev = torch.device("cuda:0")


class Smi:
    def __init__(self):
        self.lines = []
        self.p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                   "-lms", "100"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for ln in self.p.stdout:
            self.lines.append(ln)

    def mark(self):
        return len(self.lines)

    def stats(self, a, b):
        v = [tuple(float(t) for t in ln.split(",")) for ln in self.lines[a:b] if "," in ln]
        if not v:
            return (0, 0)
        v = v[len(v) // 3:]
        return (sorted(x[0] for x in v)[len(v) // 2], sum(x[1] for x in v) / len(v))


def loop(fn, seconds, flops):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return flops * n / ms / 1e9


smi = Smi()
time.sleep(0.5)
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
for (m, n, k) in [(131328, 3072, 1024), (131328, 4096, 1024), (131328, 1024, 4096), (131328, 1024, 1024), (25344, 4096, 1024)]:
    a = torch.randn(m, k, device=dev).bfloat16()
    b = torch.randn(n, k, device=dev).bfloat16()
    out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * m * n * k
    i0 = smi.mark()
    ours = loop(lambda: ops.gemm(a, b, m=m, n=n, k=k, out=out), secs, fl)
    i1 = smi.mark()
    bt = b.t()
    cub = loop(lambda: torch.matmul(a, bt, out=out), secs, fl)
    i2 = smi.mark()
    print(f"[SUSTAINED] m={m} n={n} k={k}: mb {ours:.0f} TFLOP/s (clk, W = {smi.stats(i0, i1)}) | cuBLAS {cub:.0f} TFLOP/s (clk, W = {smi.stats(i1, i2)})",
          flush=True)
smi.p.terminate()
