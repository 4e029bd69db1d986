"""GPU bring-up check for mb_gemm: run `python scripts/check_gemm.py <group>` on a B200.

Each group runs in its own process (a trapped kernel poisons the CUDA context), prints a verdict
per case and a short diagnosis of the error pattern when a case fails.
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from mirage_b200 import _lib as L  # noqa: E402
from mirage_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def report(name, got, ref, tol=2e-2):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-9
    rel = err.max().item() / scale
    bad = err > tol * scale
    ok = rel <= tol and torch.isfinite(got).all().item()
    print(f"[{'PASS' if ok else 'FAIL'}] {name}: max_rel={rel:.3e} bad_frac={bad.float().mean().item():.4f}",
          flush=True)
    if not ok:
        idx = bad.nonzero()
        if idx.numel():
            rows = idx[:, 0].unique()
            cols = idx[:, 1].unique()
            print(f"    bad rows: n={rows.numel()} first={rows[:12].tolist()}  "
                  f"bad cols: n={cols.numel()} first={cols[:12].tolist()}")
            r, c = idx[0].tolist()
            print(f"    first bad ({r},{c}): got={got[r, c].item():.5f} ref={ref[r, c].item():.5f}")
            print(f"    got[0,:8]={got[0, :8].tolist()}\n    ref[0,:8]={ref[0, :8].tolist()}")
    return ok


def mk(m, k, dtype=torch.bfloat16, scale=1.0):
    return (torch.randn(m, k, device=dev) * scale).to(dtype)


def case_fwd(m, n, k, bn=0, **kw):
    a = mk(m, k)
    w = mk(n, k, scale=k ** -0.5)
    out = ops.gemm(a, w, m=m, n=n, k=k, block_n=bn, **kw)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    return report(f"fwd KK m={m} n={n} k={k} bn={bn}", out, ref)


def group_basic():
    ok = True
    ok &= case_fwd(128, 64, 64, bn=64)
    ok &= case_fwd(128, 128, 64, bn=128)
    ok &= case_fwd(128, 256, 64, bn=256)
    ok &= case_fwd(128, 256, 256, bn=256)
    ok &= case_fwd(256, 512, 1024, bn=256)
    ok &= case_fwd(4096, 1024, 1024, bn=256)
    ok &= case_fwd(4096, 1024, 1024, bn=128)
    ok &= case_fwd(4096, 1024, 1024, bn=64)
    ok &= case_fwd(1539, 832, 256)       # ragged M and N
    ok &= case_fwd(99 * 7, 768, 768)
    ok &= case_fwd(131328, 1024, 1024)   # full cfg-2 row count, many tiles per CTA
    return ok


def group_epilogue():
    ok = True
    m, n, k = 1000, 512, 256
    a, w = mk(m, k), mk(n, k, scale=k ** -0.5)
    bias = torch.randn(n, device=dev)
    res = torch.randn(m, n, device=dev)
    base = a.float() @ w.float().t()
    out = ops.gemm(a, w, m=m, n=n, k=k, bias=bias)
    ok &= report("bias", out, base + bias)
    out = ops.gemm(a, w, m=m, n=n, k=k, bias=bias, out_dtype=torch.float32)
    ok &= report("bias f32 out", out, base + bias, tol=5e-3)
    pre = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
    out = ops.gemm(a, w, m=m, n=n, k=k, bias=bias, gelu=True, aux_out=pre)
    ok &= report("bias+gelu", out, torch.nn.functional.gelu(base + bias))
    ok &= report("gelu aux_out", pre, base + bias)
    out = ops.gemm(a, w, m=m, n=n, k=k, bias=bias, residual=res, out_dtype=torch.float32)
    ok &= report("bias+residual f32", out, base + bias + res, tol=5e-3)
    # in-place residual
    x = res.clone()
    ops.gemm(a, w, m=m, n=n, k=k, bias=bias, residual=x, out=x)
    ok &= report("bias+residual in place", x, base + bias + res, tol=5e-3)
    pos = torch.randn(250, n, device=dev)
    out = ops.gemm(a, w, m=m, n=n, k=k, bias=bias, residual=pos, res_period=250,
                   out_dtype=torch.float32)
    ok &= report("bias+periodic residual", out, base + bias + pos.repeat(4, 1), tol=5e-3)
    h = mk(m, n)
    out = ops.gemm(a, w, m=m, n=n, k=k, dgelu_aux=h)
    hf = h.float().requires_grad_(True)
    g = torch.autograd.grad(torch.nn.functional.gelu(hf).sum(), hf)[0]
    ok &= report("dgelu", out, base * g)
    # dgrad through the GELU with the bias gradient (column sums of the result) from the epilogue
    for (m2, n2, k2) in [(1000, 512, 256), (25344, 4096, 1024)]:
        dy, w2 = mk(m2, k2), mk(k2, n2, scale=k2 ** -0.5)
        h2 = mk(m2, n2)
        cs = torch.full((n2,), 3.0, device=dev)
        out2 = ops.gemm(dy, w2, m=m2, n=n2, k=k2, b_layout=L.MB_MAJOR_MN, dgelu_aux=h2, colsum_out=cs)
        hf2 = h2.float().requires_grad_(True)
        g2 = torch.autograd.grad(torch.nn.functional.gelu(hf2).sum(), hf2)[0]
        ref2 = (dy.float() @ w2.float()) * g2
        ok &= report(f"dgelu + colsum: output m={m2}", out2, ref2)
        ok &= report(f"dgelu + colsum: column sums m={m2}", cs[None], (3.0 + ref2.double().sum(0)).float()[None], tol=2e-3)
    return ok


def group_dgrad():
    ok = True
    for (m, n_out, k_in, bn) in [(128, 64, 64, 64), (128, 64, 256, 256), (256, 128, 128, 128),
                                 (1000, 768, 512, 0), (4096, 4096, 1024, 0)]:
        dy = mk(m, n_out)
        w = mk(n_out, k_in, scale=n_out ** -0.5)
        # dx[m, k_in] = dy[m, n_out] @ w[n_out, k_in]; reduction over n_out; B = w is MN-major
        out = ops.gemm(dy, w, m=m, n=k_in, k=n_out, b_layout=L.MB_MAJOR_MN, block_n=bn)
        ok &= report(f"dgrad m={m} n_out={n_out} k_in={k_in} bn={bn}", out, dy.float() @ w.float())
    return ok


def group_wgrad():
    ok = True
    for (t, n_out, k_in, bn, splits) in [(64, 128, 64, 64, 1), (64, 128, 256, 256, 1),
                                         (256, 256, 256, 128, 1), (1000, 768, 512, 0, 1),
                                         (4096, 1024, 1024, 0, 8), (25344, 1024, 4096, 0, 4)]:
        dy = mk(t, n_out, scale=t ** -0.5)
        x = mk(t, k_in)
        out = torch.zeros(n_out, k_in, device=dev)
        # dw[n_out, k_in] = dy^T @ x; reduction over tokens; both operands MN-major
        ops.gemm(dy, x, m=n_out, n=k_in, k=t, a_layout=L.MB_MAJOR_MN, b_layout=L.MB_MAJOR_MN,
                 out=out, block_n=bn, k_splits=splits, atomic=True)
        ok &= report(f"wgrad t={t} n_out={n_out} k_in={k_in} bn={bn} splits={splits}", out,
                     dy.float().t() @ x.float(), tol=1e-2)
    return ok


def group_patch():
    ok = True
    for (b, d) in [(1, 256), (3, 768), (8, 1024)]:
        img = torch.rand(b, 1, 512, 512, device=dev)
        w = torch.randn(d, 1, 32, 32, device=dev) * 0.03
        bias = torch.randn(d, device=dev) * 0.1
        pos = torch.randn(256, d, device=dev)
        out = ops.gemm(img, w.view(d, 1024), m=b * 256, n=d, k=1024, a_layout=L.MB_A_PATCH32,
                       img_hw=(512, 512), bias=bias, residual=pos, res_period=256,
                       out_dtype=torch.float32)
        ref = torch.nn.functional.conv2d(img, w, bias, stride=32).flatten(2).transpose(1, 2) + pos
        ok &= report(f"patch32 tf32 b={b} d={d}", out.view(b, 256, d).flatten(0, 1), ref.flatten(0, 1),
                     tol=5e-3)
    # plain tf32 K-major GEMM
    a = torch.randn(300, 512, device=dev)
    w = torch.randn(256, 512, device=dev) * 0.05
    out = ops.gemm(a, w, m=300, n=256, k=512, out_dtype=torch.float32)
    ok &= report("tf32 KK", out, a @ w.t(), tol=5e-3)
    return ok


def group_pair():
    """cta_group::2 kernel (forced): every layout / epilogue the 1-CTA kernel supports."""
    ok = True
    for (m, n, k, bn) in [(256, 256, 64, 256), (256, 256, 256, 256), (512, 512, 1024, 256), (256, 128, 128, 128),
                          (4096, 1024, 1024, 256), (4096, 1024, 1024, 128), (1539, 832, 256, 128),
                          (693, 768, 768, 256), (131328, 1024, 1024, 256)]:
        a, w = mk(m, k), mk(n, k, scale=k ** -0.5)
        out = ops.gemm(a, w, m=m, n=n, k=k, block_n=bn, cta_pair=2)
        ok &= report(f"pair fwd m={m} n={n} k={k} bn={bn}", out, a.float() @ w.float().t())
    m, n, k = 1000, 512, 256
    a, w = mk(m, k), mk(n, k, scale=k ** -0.5)
    bias, res = torch.randn(n, device=dev), torch.randn(m, n, device=dev)
    base = a.float() @ w.float().t()
    pre = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
    out = ops.gemm(a, w, m=m, n=n, k=k, bias=bias, gelu=True, aux_out=pre, cta_pair=2)
    ok &= report("pair bias+gelu", out, torch.nn.functional.gelu(base + bias))
    ok &= report("pair gelu aux_out", pre, base + bias)
    x = res.clone()
    ops.gemm(a, w, m=m, n=n, k=k, bias=bias, residual=x, out=x, cta_pair=2)
    ok &= report("pair bias+residual in place", x, base + bias + res, tol=5e-3)
    for (m, n_out, k_in) in [(256, 128, 256), (1000, 768, 512), (4096, 4096, 1024)]:
        dy, w = mk(m, n_out), mk(n_out, k_in, scale=n_out ** -0.5)
        out = ops.gemm(dy, w, m=m, n=k_in, k=n_out, b_layout=L.MB_MAJOR_MN, cta_pair=2)
        ok &= report(f"pair dgrad m={m} n_out={n_out} k_in={k_in}", out, dy.float() @ w.float())
    for (t, n_out, k_in, splits) in [(64, 256, 256, 1), (1000, 768, 512, 1), (4096, 1024, 1024, 8),
                                     (25344, 1024, 4096, 4)]:
        dy, x = mk(t, n_out, scale=t ** -0.5), mk(t, k_in)
        out = torch.zeros(n_out, k_in, device=dev)
        ops.gemm(dy, x, m=n_out, n=k_in, k=t, a_layout=L.MB_MAJOR_MN, b_layout=L.MB_MAJOR_MN, out=out,
                 k_splits=splits, atomic=True, cta_pair=2)
        ok &= report(f"pair wgrad t={t} n_out={n_out} k_in={k_in} splits={splits}", out,
                     dy.float().t() @ x.float(), tol=1e-2)
    for (b, d) in [(2, 256), (8, 1024)]:
        img = torch.rand(b, 1, 512, 512, device=dev)
        w = torch.randn(d, 1, 32, 32, device=dev) * 0.03
        bias = torch.randn(d, device=dev) * 0.1
        pos = torch.randn(256, d, device=dev)
        out = ops.gemm(img, w.view(d, 1024), m=b * 256, n=d, k=1024, a_layout=L.MB_A_PATCH32,
                       img_hw=(512, 512), bias=bias, residual=pos, res_period=256,
                       out_dtype=torch.float32, cta_pair=2)
        ref = torch.nn.functional.conv2d(img, w, bias, stride=32).flatten(2).transpose(1, 2) + pos
        ok &= report(f"pair patch32 tf32 b={b} d={d}", out, ref.flatten(0, 1), tol=5e-3)
    return ok


def group_perf():
    shapes = [(131328, 3072, 1024), (131328, 1024, 1024), (131328, 4096, 1024), (131328, 1024, 4096),
              (25344, 4096, 1024), (8192, 8192, 8192)]
    for (m, n, k) in shapes:
        a, w = mk(m, k), mk(n, k, scale=k ** -0.5)
        bias = torch.randn(n, device=dev)
        out = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        res = {}
        for mode in (1, 2):
            for _ in range(3):
                ops.gemm(a, w, m=m, n=n, k=k, bias=bias, out=out, cta_pair=mode)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                ops.gemm(a, w, m=m, n=n, k=k, bias=bias, out=out, cta_pair=mode)
            e1.record()
            torch.cuda.synchronize()
            res[mode] = e0.elapsed_time(e1) / iters
        ms = res[2]
        tf = 2.0 * m * n * k / ms / 1e9
        for _ in range(3):
            torch.nn.functional.linear(a, w)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            torch.nn.functional.linear(a, w)
        e1.record()
        torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1) / iters
        print(f"[PERF] m={m} n={n} k={k}: 1-CTA {res[1]:.3f} ms = {2.0 * m * n * k / res[1] / 1e9:.0f} TFLOP/s | "
              f"pair {ms:.3f} ms = {tf:.0f} TFLOP/s | "
              f"cuBLAS {ms_t:.3f} ms = {2.0 * m * n * k / ms_t / 1e9:.0f} TFLOP/s", flush=True)
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


def group_maps():
    """EPI_MAP: output row map, periodic residual, un-patchify -- specialised epilogue vs torch."""
    ok = True
    # un-patchify: tokens [B*gh*gw, C*ph*pw] -> image [B, C, gh*ph, gw*pw]
    for (B, C_, ph, pw, gh, gw, k) in [(3, 1, 32, 32, 16, 16, 256), (2, 13, 8, 8, 16, 16, 256), (1, 1, 32, 32, 16, 16, 64)]:
        m, n = B * gh * gw, C_ * ph * pw
        a, w = mk(m, k), mk(n, k, scale=k ** -0.5)
        bias = torch.randn(n, device=dev)
        img = ops.gemm(a, w, m=m, n=n, k=k, bias=bias, out_dtype=torch.float32, unpatch=(C_, ph, pw, gh, gw))
        ref = (a.float() @ w.float().t() + bias).reshape(B, gh, gw, C_, ph, pw).permute(0, 3, 1, 4, 2, 5)
        ok &= report(f"unpatch B={B} C={C_} p={ph}", img.reshape(B * C_ * gh * ph, gw * pw),
                     ref.reshape(B * C_ * gh * ph, gw * pw), tol=5e-3)
    # row map + periodic residual: tokens of one modality written into a [B, 513, D] buffer at offset 256
    B, ntok, d, k = 5, 256, 512, 256
    a, w = mk(B * ntok, k), mk(d, k, scale=k ** -0.5)
    bias, pos = torch.randn(d, device=dev), torch.randn(ntok, d, device=dev)
    buf = torch.full((B, 513, d), 7.0, device=dev)
    ops.gemm(a, w, m=B * ntok, n=d, k=k, bias=bias, residual=pos, res_period=ntok, out=buf.view(B * 513, d),
             out_row_map=(ntok, 513, 256))
    ref = (a.float() @ w.float().t() + bias).view(B, ntok, d) + pos
    ok &= report("row map + periodic residual", buf[:, 256:512].reshape(-1, d), ref.reshape(-1, d), tol=5e-3)
    untouched = bool((buf[:, :256] == 7.0).all() and (buf[:, 512:] == 7.0).all())
    print(f"[{'PASS' if untouched else 'FAIL'}] row map leaves the other rows untouched", flush=True)
    return ok and untouched


def group_perfpre():
    """cfg-4 sized GEMMs (25344 token rows): 256- vs 128-wide CTA-pair tiles (wave quantisation)."""
    m = 25344
    for (n, k, kind) in [(1024, 1024, "res"), (1024, 4096, "res"), (1024, 4096, "dgrad"), (1024, 3072, "dgrad"),
                         (3072, 1024, "bf16"), (4096, 1024, "gelu")]:
        for bn in (256, 128):
            a = mk(m, k)
            bias = torch.randn(n, device=dev)
            if kind == "dgrad":
                w = mk(k, n, scale=k ** -0.5)
                fn = lambda: ops.gemm(a, w, m=m, n=n, k=k, b_layout=L.MB_MAJOR_MN, block_n=bn, cta_pair=2)
            elif kind == "res":
                w = mk(n, k, scale=k ** -0.5)
                x = torch.randn(m, n, device=dev)
                fn = lambda: ops.gemm(a, w, m=m, n=n, k=k, bias=bias, residual=x, out=x, block_n=bn, cta_pair=2)
            elif kind == "gelu":
                w = mk(n, k, scale=k ** -0.5)
                pre = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
                fn = lambda: ops.gemm(a, w, m=m, n=n, k=k, bias=bias, gelu=True, aux_out=pre, block_n=bn, cta_pair=2)
            else:
                w = mk(n, k, scale=k ** -0.5)
                fn = lambda: ops.gemm(a, w, m=m, n=n, k=k, bias=bias, block_n=bn, cta_pair=2)
            ms = _time(fn, iters=20)
            print(f"[PERF] {kind:6s} m={m} n={n} k={k} pair bn={bn}: {ms * 1e3:.1f} us = {2.0 * m * n * k / ms / 1e9:.0f} TFLOP/s",
                  flush=True)
    return True


def group_perfepi():
    """Epilogue variants of the encoder's GEMMs at cfg-2 size (the tile loop is identical; only the
    epilogue differs), against the plain bf16 store."""
    m = 131328
    for (n, k, what) in [(4096, 1024, "gelu"), (4096, 1024, "gelu+aux"), (1024, 4096, "res"), (1024, 1024, "res"),
                         (3072, 1024, "bf16")]:
        a, w = mk(m, k), mk(n, k, scale=k ** -0.5)
        bias = torch.randn(n, device=dev)
        if what == "res":
            x = torch.randn(m, n, device=dev)
            fn = lambda: ops.gemm(a, w, m=m, n=n, k=k, bias=bias, residual=x, out=x)
        elif what.startswith("gelu"):
            out = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
            pre = torch.empty(m, n, dtype=torch.bfloat16, device=dev) if what == "gelu+aux" else None
            fn = lambda: ops.gemm(a, w, m=m, n=n, k=k, bias=bias, gelu=True, aux_out=pre, out=out)
        else:
            out = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
            fn = lambda: ops.gemm(a, w, m=m, n=n, k=k, bias=bias, out=out)
        ms = _time(fn)
        print(f"[PERF] {what:9s} m={m} n={n} k={k}: {ms:.3f} ms = {2.0 * m * n * k / ms / 1e9:.0f} TFLOP/s", flush=True)
    # patch embedding (tf32, image read by TMA boxes) and out_proj + un-patchify at cfg-4 size
    img = torch.rand(256, 1, 512, 512, device=dev)
    w = torch.randn(1024, 1024, device=dev) * 0.03
    bias, pos = torch.randn(1024, device=dev), torch.randn(256, 1024, device=dev)
    out = torch.empty(256 * 256, 1024, device=dev)
    ms = _time(lambda: ops.gemm(img, w, m=65536, n=1024, k=1024, a_layout=L.MB_A_PATCH32, img_hw=(512, 512),
                                bias=bias, residual=pos, res_period=256, out=out))
    print(f"[PERF] patch-embed tf32 65536x1024x1024 +bias+pos: {ms:.3f} ms = {2.0 * 65536 * 1024 * 1024 / ms / 1e9:.0f} TFLOP/s",
          flush=True)
    a, w2 = mk(65536, 256), mk(1024, 256, scale=1 / 16)
    ms = _time(lambda: ops.gemm(a, w2, m=65536, n=1024, k=256, bias=bias, out_dtype=torch.float32,
                                unpatch=(1, 32, 32, 16, 16)))
    print(f"[PERF] out_proj+unpatch 65536x1024x256: {ms:.3f} ms = {65536 * 1024 * 4 / ms / 1e6:.0f} GB/s written", flush=True)
    return True


if __name__ == "__main__":
    grp = sys.argv[1]
    t0 = time.time()
    print(f"=== group {grp} on {torch.cuda.get_device_name(0)} ===", flush=True)
    ok = globals()[f"group_{grp}"]()
    torch.cuda.synchronize()
    print(f"=== group {grp}: {'ALL PASS' if ok else 'FAILURES'} ({time.time() - t0:.1f}s) ===", flush=True)
    sys.exit(0 if ok else 1)
