"""Debug: per-event SM-clock timeline of attention4 (build with MB_NVCC_EXTRA=-DMB_ATTN4_TIMING)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from mirage_b200 import _lib as L
from mirage_b200 import ops

dev = torch.device("cuda:0")
B, H, n, hd = 256, 16, 513, 64
D = H * hd
qkv = torch.randn(B * n, 3 * D, device=dev).bfloat16()
out = torch.empty(B * n, D, dtype=torch.bfloat16, device=dev)
for _ in range(2):
    ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], batch=B, heads=H, nq=n, nk=n, head_dim=hd, scale=hd ** -0.5, out=out)
torch.cuda.synchronize()
TN, G, Q, E = 6, 4, 4, 12
buf = np.zeros((TN, G, Q, E), dtype=np.int64)
lib = L.lib()
lib.mb_debug_attn4_timing.restype = C.c_int
assert lib.mb_debug_attn4_timing(buf.ctypes.data_as(C.c_void_p)) == 0
t0 = buf[buf > 0].min()
names = {0: "arrive p_full", 1: "issuer: p_full seen", 2: "issuer: operands ready, waiting for P", 3: "issuer: PV committed",
         4: "issuer: k_full seen", 5: "issuer: S committed", 6: "softmax: s_full seen", 7: "softmax: S loaded",
         8: "softmax: max done", 9: "softmax: exp done"}
for g in range(G):
    print(f"--- tile {g}")
    ev = []
    for b in range(TN):
        for q in range(Q):
            for e in range(E):
                if buf[b, g, q, e] > 0:
                    ev.append((int(buf[b, g, q, e] - t0), b, q, e))
    ev.sort()
    for t, b, q, e in ev:
        if e in (0, 6) or q == 0:
            print(f"  {t:7d}  blk {b} q{q}  {names.get(e, e)}")
