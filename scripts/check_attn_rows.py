"""GPU bring-up check for mb_attn_fwd and the row kernels: `python scripts/check_attn_rows.py <group>`."""
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from mirage_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def report(name, got, ref, tol=2e-2):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-9
    rel = err.max().item() / scale
    ok = rel <= tol and torch.isfinite(got).all().item()
    print(f"[{'PASS' if ok else 'FAIL'}] {name}: max_rel={rel:.3e} "
          f"bad_frac={(err > tol * scale).float().mean().item():.4f}", flush=True)
    if not ok:
        bad = (err > tol * scale) | ~torch.isfinite(got)
        idx = bad.reshape(bad.shape[0], -1).any(dim=1).nonzero().flatten()
        print(f"    bad rows n={idx.numel()} first={idx[:16].tolist()} last={idx[-4:].tolist()}")
        print(f"    got[0,:6]={got.reshape(got.shape[0], -1)[0, :6].tolist()}\n"
              f"    ref[0,:6]={ref.reshape(ref.shape[0], -1)[0, :6].tolist()}")
    return ok


def attn_case(B, H, nq, nk, hd, self_attn=True, amp=1.0):
    D = H * hd
    scale = hd ** -0.5
    if self_attn:
        qkv = (torch.randn(B * nq, 3 * D, device=dev) * amp).bfloat16()
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    else:
        q = torch.randn(B * nq, D, device=dev).bfloat16()
        kv = torch.randn(B * nk, 2 * D, device=dev).bfloat16()
        k, v = kv[:, :D], kv[:, D:]
    lse = torch.empty(B, H, nq, device=dev)
    out = ops.attention(q, k, v, batch=B, heads=H, nq=nq, nk=nk, head_dim=hd, scale=scale, lse=lse)
    torch.cuda.synchronize()
    qf = q.float().reshape(B, nq, H, hd).transpose(1, 2)
    kf = k.float().reshape(B, nk, H, hd).transpose(1, 2)
    vf = v.float().reshape(B, nk, H, hd).transpose(1, 2)
    s = (qf @ kf.transpose(-1, -2)) * scale
    ref = (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(B * nq, D)
    ok = report(f"attn B={B} H={H} nq={nq} nk={nk} hd={hd} amp={amp}", out, ref)
    ok &= report("   lse", lse, torch.logsumexp(s, -1), tol=1e-3)
    return ok


def group_attn64():
    ok = True
    ok &= attn_case(1, 1, 128, 128, 64)
    ok &= attn_case(1, 1, 128, 256, 64, self_attn=False)
    ok &= attn_case(1, 2, 256, 256, 64)
    ok &= attn_case(2, 3, 99, 99, 64)
    ok &= attn_case(2, 12, 513, 513, 64)
    ok &= attn_case(3, 16, 257, 257, 64)
    ok &= attn_case(2, 4, 1025, 1025, 64)
    ok &= attn_case(2, 4, 200, 77, 64, self_attn=False)
    # peeled remainders (global tokens appended last): key tail inside the tile kernel, query tail rows
    ok &= attn_case(2, 4, 258, 258, 64)                    # 2 global tokens
    ok &= attn_case(1, 3, 772, 772, 64)                    # 3 modalities + 4 global tokens
    ok &= attn_case(2, 2, 256, 513, 64, self_attn=False)   # key tail only
    ok &= attn_case(2, 2, 513, 256, 64, self_attn=False)   # query tail only
    ok &= attn_case(2, 2, 129, 129, 64)                    # query tail peeled, keys padded (nk < 256)
    ok &= attn_case(2, 2, 133, 261, 64, self_attn=False)   # remainder 5: nothing peeled
    # many work items per persistent CTA (cross-item pipeline, ring release by the epilogue warps)
    ok &= attn_case(24, 16, 513, 513, 64)
    ok &= attn_case(40, 16, 257, 257, 64)
    ok &= attn_case(64, 16, 99, 99, 64)
    ok &= attn_case(20, 16, 514, 514, 64)                  # two peeled rows: separate tail kernel
    # large score range: the lazy running-max rescale of O in TMEM fires (four-tile kernel and two-tile kernel)
    ok &= attn_case(3, 4, 513, 513, 64, amp=3.0)
    ok &= attn_case(2, 4, 1025, 1025, 64, amp=4.0)
    ok &= attn_case(3, 4, 257, 257, 64, amp=3.0)
    ok &= attn_case(2, 2, 2049, 2049, 64)
    ok &= attn_case(300, 16, 513, 513, 64)                 # > 2 waves of (batch, head) items per CTA
    # enough (batch, head-group) CTAs for the four-heads-per-CTA tail kernel
    ok &= attn_case(160, 16, 257, 257, 64)
    ok &= attn_case(200, 12, 130, 130, 64)
    # single-tile kernel (attention_small.cu): nq, nk <= 128 -- ragged sizes, cross shapes, one key, and enough
    # (batch, head) problems for every slot of a persistent CTA to be reused several times
    ok &= attn_case(1, 1, 1, 1, 64)
    ok &= attn_case(2, 2, 17, 64, 64, self_attn=False)
    ok &= attn_case(2, 2, 100, 33, 64, self_attn=False)
    ok &= attn_case(3, 5, 65, 65, 64)
    ok &= attn_case(2, 4, 128, 97, 64, self_attn=False)
    ok &= attn_case(5, 16, 99, 99, 64, amp=4.0)
    ok &= attn_case(256, 16, 99, 99, 64)
    ok &= attn_case(97, 12, 99, 99, 64)
    return ok


def group_attn32():
    ok = True
    ok &= attn_case(1, 1, 128, 128, 32)
    ok &= attn_case(2, 8, 256, 256, 32)
    ok &= attn_case(2, 8, 256, 99, 32, self_attn=False)
    return ok


def group_rowops():
    ok = True
    for (rows, D) in [(1000, 1024), (513, 768), (77, 256), (8, 128)]:
        x = torch.randn(rows, D, device=dev) * 2 + 0.5
        w = torch.randn(D, device=dev)
        b = torch.randn(D, device=dev)
        y, mean, rstd = ops.layernorm(x, w, b, save_stats=True)
        ref = F.layer_norm(x, (D,), w, b, 1e-6)
        ok &= report(f"layernorm bf16 rows={rows} D={D}", y, ref, tol=1e-2)
        y32 = ops.layernorm(x, w, b, out_dtype=torch.float32)
        ok &= report(f"layernorm f32 rows={rows} D={D}", y32, ref, tol=1e-5)
        # backward
        dy = torch.randn(rows, D, device=dev)
        dres = torch.randn(rows, D, device=dev)
        xr = x.clone().requires_grad_(True)
        wr = w.clone().requires_grad_(True)
        br = b.clone().requires_grad_(True)
        F.layer_norm(xr, (D,), wr, br, 1e-6).backward(dy)
        dx, dw, db = ops.layernorm_bwd(dy, x, w, mean, rstd, dres=dres)
        ok &= report("   ln bwd dx(+dres)", dx, xr.grad + dres, tol=1e-4)
        ok &= report("   ln bwd dw", dw, wr.grad, tol=1e-4)
        ok &= report("   ln bwd db", db, br.grad, tol=1e-4)
        dx2, dx2b, _, _ = ops.layernorm_bwd(dy.bfloat16(), x, w, mean, rstd, want_bf16=True)
        ok &= report("   ln bwd dx (bf16 dy)", dx2, xr.grad, tol=1e-2)
        exact = torch.equal(dx2b, dx2.bfloat16())
        print(f"[{'PASS' if exact else 'FAIL'}]    ln bwd bf16 twin == round(dx)", flush=True)
        ok &= exact
    for (rows, cols) in [(5000, 768), (25344, 1024), (25344, 3072), (65536, 256), (999, 832), (3, 8), (70, 104)]:
        a = torch.randn(rows, cols, device=dev)
        ok &= report(f"colsum f32 {rows}x{cols}", ops.colsum(a)[None], a.double().sum(0).float()[None], tol=1e-4)
        ab = a.bfloat16()
        ok &= report(f"colsum bf16 {rows}x{cols}", ops.colsum(ab)[None], ab.double().sum(0).float()[None], tol=1e-4)
    acc = torch.ones(1024, device=dev)
    a = torch.randn(25344, 1024, device=dev).bfloat16()
    ops.colsum(a, into=acc)
    ops.colsum(a, into=acc)
    ok &= report("colsum accumulate (single-kernel vector reductions)", acc[None],
                 (1 + 2 * a.double().sum(0)).float()[None], tol=1e-4)
    big = torch.randn(4096, 3072, device=dev).bfloat16()
    ok &= report("colsum bf16 column slice (ld 3072)", ops.colsum(big[:, 1024:2048])[None],
                 big[:, 1024:2048].double().sum(0).float()[None], tol=1e-4)
    # gather / scatter: bit exact
    B, n_src, D, n_keep = 5, 768, 1024, 98
    src = torch.randn(B, n_src, D, device=dev)
    ids = torch.stack([torch.randperm(n_src, device=dev)[:n_keep] for _ in range(B)])
    glob = torch.randn(1, D, device=dev)
    out = ops.token_gather(src, ids, glob)
    ref = torch.cat([torch.gather(src, 1, ids[..., None].expand(-1, -1, D)), glob.expand(B, 1, D)], 1)
    exact = torch.equal(out, ref)
    print(f"[{'PASS' if exact else 'FAIL'}] token_gather bit-exact", flush=True)
    ok &= exact
    dout = torch.randn(B, n_keep + 1, D, device=dev)
    dsrc, dglob = ops.token_gather_bwd(dout, ids, n_src, 1)
    ref_d = torch.zeros_like(src).scatter_(1, ids[..., None].expand(-1, -1, D), dout[:, :n_keep])
    exact = torch.equal(dsrc, ref_d)
    print(f"[{'PASS' if exact else 'FAIL'}] token_scatter bit-exact", flush=True)
    ok &= exact
    ok &= report("global token grad", dglob, dout[:, n_keep:].sum(0), tol=1e-5)
    buf = torch.zeros(B, 513, D, device=dev)
    ops.fill_global_rows(glob, buf, 512)
    ok &= bool(torch.equal(buf[:, 512], glob.expand(B, D))) and bool((buf[:, :512] == 0).all())
    x = torch.randn(64, 256, device=dev)
    ok &= bool(torch.equal(ops.cast_bf16(x), x.bfloat16()))
    print(f"[{'PASS' if ok else 'FAIL'}] fill/cast", flush=True)
    return ok


def group_attnperf():
    for (B, H, n, hd) in [(256, 16, 513, 64), (256, 12, 513, 64), (256, 16, 99, 64), (64, 16, 257, 64)]:
        D = H * hd
        qkv = torch.randn(B * n, 3 * D, device=dev).bfloat16()
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
        out = torch.empty(B * n, D, dtype=torch.bfloat16, device=dev)
        for _ in range(3):
            ops.attention(q, k, v, batch=B, heads=H, nq=n, nk=n, head_dim=hd, scale=hd ** -0.5, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            ops.attention(q, k, v, batch=B, heads=H, nq=n, nk=n, head_dim=hd, scale=hd ** -0.5, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 4.0 * B * H * n * n * hd
        q4 = qkv.view(B, n, 3, H, hd).permute(2, 0, 3, 1, 4)
        for _ in range(3):
            F.scaled_dot_product_attention(q4[0], q4[1], q4[2])
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            F.scaled_dot_product_attention(q4[0], q4[1], q4[2])
        e1.record()
        torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1) / iters
        print(f"[PERF] attn B={B} H={H} n={n} hd={hd}: mb {ms:.3f} ms = {fl / ms / 1e9:.0f} TFLOP/s | "
              f"torch SDPA {ms_t:.3f} ms = {fl / ms_t / 1e9:.0f} TFLOP/s", flush=True)
    # LayerNorm bandwidth
    x = torch.randn(131328, 1024, device=dev)
    w = torch.ones(1024, device=dev)
    b = torch.zeros(1024, device=dev)
    for _ in range(3):
        ops.layernorm(x, w, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.layernorm(x, w, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"[PERF] layernorm 131328x1024 f32->bf16: {ms:.3f} ms = {x.numel() * 6 / ms / 1e6:.0f} GB/s", flush=True)
    return True


def _time(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def group_rowperf():
    """HBM-bound row kernels at pretraining size (cfg 4: 256 x 99 token rows, D = 1024)."""
    rows, D = 25344, 1024
    x = torch.randn(rows, D, device=dev)
    w, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    y, mean, rstd = ops.layernorm(x, w, b, save_stats=True)
    dy = torch.randn(rows, D, device=dev).bfloat16()
    dres = torch.randn(rows, D, device=dev)
    ms = _time(lambda: ops.layernorm_bwd(dy, x, w, mean, rstd, dres=dres, want_bf16=True))
    print(f"[PERF] layernorm_bwd {rows}x{D} (+dres, +bf16 twin): {ms * 1e3:.1f} us = "
          f"{rows * D * 16 / ms / 1e6:.0f} GB/s", flush=True)
    ms = _time(lambda: ops.layernorm(x, w, b, save_stats=True))
    print(f"[PERF] layernorm_fwd {rows}x{D}: {ms * 1e3:.1f} us = {rows * D * 6 / ms / 1e6:.0f} GB/s", flush=True)
    for cols in (1024, 3072, 4096):
        a = torch.randn(rows, cols, device=dev).bfloat16()
        ms = _time(lambda: ops.colsum(a))
        print(f"[PERF] colsum bf16 {rows}x{cols}: {ms * 1e3:.1f} us = {rows * cols * 2 / ms / 1e6:.0f} GB/s",
              flush=True)
    a = torch.randn(65536, 256, device=dev).bfloat16()
    ms = _time(lambda: ops.colsum(a))
    print(f"[PERF] colsum bf16 65536x256: {ms * 1e3:.1f} us = {65536 * 256 * 2 / ms / 1e6:.0f} GB/s", flush=True)
    return True


if __name__ == "__main__":
    grp = sys.argv[1]
    t0 = time.time()
    print(f"=== group {grp} on {torch.cuda.get_device_name(0)} ===", flush=True)
    ok = globals()[f"group_{grp}"]()
    torch.cuda.synchronize()
    print(f"=== group {grp}: {'ALL PASS' if ok else 'FAILURES'} ({time.time() - t0:.1f}s) ===", flush=True)
    sys.exit(0 if ok else 1)
