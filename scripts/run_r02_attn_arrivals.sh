python scripts/prof_attn_bwd_shapes.py
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x 2>&1 | tail -3
python scripts/prof_rowkernels.py 2>&1 | grep attn
python - <<'P'
import sys, torch
sys.path.insert(0, '.')
from mirage_b200 import ops
dev = torch.device('cuda:0')
for (B, H, N, hd) in ((64, 16, 257, 64), (256, 8, 256, 32)):
    D = H * hd
    qkv = torch.randn(B * N, 3 * D, device=dev).bfloat16()
    f = lambda: ops.attention(qkv[:, :D], qkv[:, D:2*D], qkv[:, 2*D:], batch=B, heads=H, nq=N, nk=N, head_dim=hd, scale=hd ** -0.5)
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"attn_fwd two-tile B={B} H={H} N={N} hd={hd}: {us:.1f} us {4.0*B*H*N*N*hd/us/1e6:.0f} TFLOP/s")
P
for wl in pretrain_large cls_large; do python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; print('$wl', d['value'], d['ms_per_step'], {n:k[n] for n in ('attn_bwd','attn_fwd')})"; done
