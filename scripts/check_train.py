"""GPU bring-up check for the training-path kernels: `python scripts/check_train.py <group>`."""
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from mirage_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def report(name, got, ref, tol=2e-2):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-9
    rel = err.max().item() / scale
    fro = (got - ref).norm().item() / (ref.norm().item() + 1e-12)
    ok = rel <= tol and torch.isfinite(got).all().item()
    print(f"[{'PASS' if ok else 'FAIL'}] {name}: max_rel={rel:.3e} rel_fro={fro:.3e}", flush=True)
    if not ok:
        bad = ((err > tol * scale) | ~torch.isfinite(got)).reshape(got.shape[0], -1).any(1).nonzero().flatten()
        print(f"    bad rows n={bad.numel()} first={bad[:12].tolist()}")
        print(f"    got[0,:6]={got.reshape(got.shape[0], -1)[0, :6].tolist()}\n"
              f"    ref[0,:6]={ref.reshape(ref.shape[0], -1)[0, :6].tolist()}")
    return ok


def attn_bwd_case(B, H, nq, nk, hd, self_attn=True):
    D = H * hd
    scale = hd ** -0.5
    if self_attn:
        qkv = torch.randn(B * nq, 3 * D, device=dev).bfloat16()
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    else:
        q = torch.randn(B * nq, D, device=dev).bfloat16()
        kv = torch.randn(B * nk, 2 * D, device=dev).bfloat16()
        k, v = kv[:, :D], kv[:, D:]
    lse = torch.empty(B, H, nq, device=dev)
    o = ops.attention(q, k, v, batch=B, heads=H, nq=nq, nk=nk, head_dim=hd, scale=scale, lse=lse)
    do = torch.randn(B * nq, D, device=dev).bfloat16()
    dqkv = torch.empty(B * nq, D, dtype=torch.bfloat16, device=dev)
    dk = torch.empty(B * nk, D, dtype=torch.bfloat16, device=dev)
    dv = torch.empty(B * nk, D, dtype=torch.bfloat16, device=dev)
    ops.attention_bwd(q, k, v, o, do, lse, dqkv, dk, dv, batch=B, heads=H, nq=nq, nk=nk, head_dim=hd, scale=scale)
    torch.cuda.synchronize()
    qf = q.float().reshape(B, nq, H, hd).transpose(1, 2).requires_grad_(True)
    kf = k.float().reshape(B, nk, H, hd).transpose(1, 2).requires_grad_(True)
    vf = v.float().reshape(B, nk, H, hd).transpose(1, 2).requires_grad_(True)
    ref = torch.softmax((qf @ kf.transpose(-1, -2)) * scale, -1) @ vf
    ref.backward(do.float().reshape(B, nq, H, hd).transpose(1, 2))
    tag = f"attn_bwd B={B} H={H} nq={nq} nk={nk} hd={hd}"
    ok = report(tag + " dq", dqkv, qf.grad.transpose(1, 2).reshape(B * nq, D))
    ok &= report(tag + " dk", dk, kf.grad.transpose(1, 2).reshape(B * nk, D))
    ok &= report(tag + " dv", dv, vf.grad.transpose(1, 2).reshape(B * nk, D))
    return ok


def group_attnbwd():
    ok = True
    ok &= attn_bwd_case(1, 1, 128, 128, 64)
    ok &= attn_bwd_case(2, 3, 99, 99, 64)
    ok &= attn_bwd_case(1, 2, 256, 128, 64, self_attn=False)
    ok &= attn_bwd_case(2, 4, 257, 257, 64)
    ok &= attn_bwd_case(1, 2, 513, 513, 64)
    ok &= attn_bwd_case(1, 1, 128, 128, 32)
    ok &= attn_bwd_case(2, 8, 256, 256, 32)
    ok &= attn_bwd_case(2, 8, 256, 99, 32, self_attn=False)
    # many items per persistent CTA: the cross-item software pipeline (phases, slot reuse)
    junk = torch.full((1 << 27,), float("nan"), device=dev)  # poison recycled memory
    del junk
    ok &= attn_bwd_case(40, 16, 99, 99, 64)
    ok &= attn_bwd_case(12, 16, 257, 257, 64)
    ok &= attn_bwd_case(48, 8, 256, 99, 32, self_attn=False)
    ok &= attn_bwd_case(40, 8, 256, 256, 32)
    ok &= attn_bwd_case(3, 50, 130, 70, 64, self_attn=False)
    return ok


def group_attnbwdperf():
    for (B, H, nq, nk, hd) in [(256, 16, 99, 99, 64), (256, 8, 256, 99, 32), (256, 8, 256, 256, 32), (64, 16, 257, 257, 64)]:
        D = H * hd
        q = torch.randn(B * nq, D, device=dev).bfloat16()
        kv = torch.randn(B * nk, 2 * D, device=dev).bfloat16()
        k, v = kv[:, :D], kv[:, D:]
        lse = torch.empty(B, H, nq, device=dev)
        o = ops.attention(q, k, v, batch=B, heads=H, nq=nq, nk=nk, head_dim=hd, scale=hd ** -0.5, lse=lse)
        do = torch.randn(B * nq, D, device=dev).bfloat16()
        dq, dk, dv = torch.empty_like(q), torch.empty(B * nk, D, dtype=torch.bfloat16, device=dev), torch.empty(B * nk, D, dtype=torch.bfloat16, device=dev)
        fn = lambda: ops.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, batch=B, heads=H, nq=nq, nk=nk, head_dim=hd, scale=hd ** -0.5)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"[PERF] attn_bwd B={B} H={H} nq={nq} nk={nk} hd={hd}: {ms * 1e3:.1f} us = "
              f"{10.0 * B * H * nq * nk * hd / ms / 1e9:.0f} TFLOP/s", flush=True)
    return True


def group_adapters():
    ok = True
    # semseg patches
    B, H, W, P, E, ncls = 3, 128, 128, 8, 64, 13
    labels = torch.randint(0, ncls, (B, H, W), device=dev)
    table = torch.randn(ncls, E, device=dev).bfloat16()
    a = ops.semseg_patches(labels, table, P, P)
    emb = table[labels].permute(0, 3, 1, 2)                                  # [B, E, H, W]
    ref = emb.reshape(B, E, H // P, P, W // P, P).permute(0, 2, 4, 1, 3, 5).reshape(B * 256, E * P * P)
    exact = torch.equal(a, ref)
    print(f"[{'PASS' if exact else 'FAIL'}] semseg_patches bit-exact", flush=True)
    ok &= exact
    dA = torch.randn_like(a.float()).bfloat16()
    dE = ops.class_emb_grad(labels, dA, ncls, E, P, P)
    t = table.float().requires_grad_(True)
    emb2 = t[labels].permute(0, 3, 1, 2).reshape(B, E, H // P, P, W // P, P).permute(0, 2, 4, 1, 3, 5)
    emb2.reshape(B * 256, E * P * P).backward(dA.float())
    ok &= report("class_emb_grad", dE, t.grad, tol=1e-3)
    # decoder assemble
    B, n_vis, n_glob, n_all, Dd = 4, 98, 1, 768, 256
    ctx = torch.randn(B, n_vis + n_glob, Dd, device=dev)
    mask_tok = torch.randn(Dd, device=dev)
    embt = torch.randn(n_all, Dd, device=dev)
    shuffle = torch.stack([torch.randperm(n_all, device=dev) for _ in range(B)])
    restore = torch.argsort(shuffle, dim=1)
    keep = shuffle[:, :n_vis].contiguous()
    for q_start in (0, 256, 512):
        q, c = ops.dec_assemble(ctx, mask_tok, embt, keep, restore, q_start, 256, n_glob)
        full = torch.cat([ctx[:, :n_vis], mask_tok.expand(B, n_all - n_vis, Dd)], 1)
        full = torch.gather(full, 1, restore[..., None].expand(-1, -1, Dd)) + embt
        q_ref = full[:, q_start:q_start + 256]
        c_ref = torch.cat([torch.gather(full, 1, keep[..., None].expand(-1, -1, Dd)), ctx[:, n_vis:]], 1)
        e1, e2 = torch.equal(q, q_ref), torch.equal(c, c_ref)
        print(f"[{'PASS' if e1 and e2 else 'FAIL'}] dec_assemble q_start={q_start} bit-exact q={e1} c={e2}", flush=True)
        ok &= e1 and e2
        # backward
        ctx_r = ctx.clone().requires_grad_(True)
        mt_r = mask_tok.clone().requires_grad_(True)
        em_r = embt.clone().requires_grad_(True)
        full = torch.cat([ctx_r[:, :n_vis], mt_r.expand(B, n_all - n_vis, Dd)], 1)
        full = torch.gather(full, 1, restore[..., None].expand(-1, -1, Dd)) + em_r
        qr = full[:, q_start:q_start + 256]
        cr = torch.cat([torch.gather(full, 1, keep[..., None].expand(-1, -1, Dd)), ctx_r[:, n_vis:]], 1)
        dq, dc = torch.randn_like(qr), torch.randn_like(cr)
        (qr * dq).sum().add((cr * dc).sum()).backward()
        dctx, demb, dmask = ops.dec_assemble_bwd(dq, dc, keep, restore, q_start, n_glob)
        ok &= report("   dctx", dctx.flatten(0, 1), ctx_r.grad.flatten(0, 1), tol=1e-5)
        ok &= report("   demb", demb, em_r.grad, tol=1e-5)
        ok &= report("   dmask_token", dmask[None], mt_r.grad[None], tol=1e-5)
    # unpatchify epilogue + patchify_cast
    for (C, P, T, Dd) in [(1, 32, 3 * 256, 256), (13, 8, 2 * 256, 256)]:
        x = torch.randn(T, Dd, device=dev).bfloat16()
        w = (torch.randn(C * P * P, Dd, device=dev) * Dd ** -0.5).bfloat16()
        bias = torch.randn(C * P * P, device=dev)
        img = ops.gemm(x, w, m=T, n=C * P * P, k=Dd, bias=bias, out_dtype=torch.float32, unpatch=(C, P, P, 16, 16))
        y = x.float() @ w.float().t() + bias
        Bn = T // 256
        ref = y.reshape(Bn, 16, 16, C, P, P).permute(0, 3, 1, 4, 2, 5).reshape(Bn, C, 16 * P, 16 * P)
        ok &= report(f"gemm unpatch C={C} P={P}", img.flatten(0, 1), ref.flatten(0, 1), tol=5e-3)
        tok = ops.patchify_cast(ref.contiguous(), P, P)
        ok &= report(f"patchify_cast C={C} P={P}", tok, y, tol=1e-2)
    return ok


def group_loss():
    from oracle import mirage_oracle as O
    from mirage_b200.criterion import MaskedCrossEntropyLoss, MaskedMSELoss
    ok = True
    g = torch.Generator().manual_seed(5)
    B = 3
    pred = torch.randn(B, 1, 512, 512, generator=g)
    tgt = torch.rand(B, 1, 512, 512, generator=g)
    logits = torch.randn(B, 13, 128, 128, generator=g)
    labels = torch.randint(0, 13, (B, 128, 128), generator=g)
    masks = {"random": (torch.rand(B, 256, generator=g) > 0.3).long(),
             "all_masked": torch.ones(B, 256, dtype=torch.long),
             "none_masked": torch.zeros(B, 256, dtype=torch.long),
             "one_empty_sample": torch.cat([torch.zeros(1, 256, dtype=torch.long),
                                            (torch.rand(B - 1, 256, generator=g) > 0.5).long()]),
             "no_mask": None}
    mse, ce, ce_ls = MaskedMSELoss((32, 32)), MaskedCrossEntropyLoss((8, 8)), MaskedCrossEntropyLoss((8, 8), label_smoothing=0.1)
    for name, m in masks.items():
        md = None if m is None else m.to(dev)
        p = pred.to(dev).requires_grad_(True)
        v = mse(p, tgt.to(dev), mask=md)
        pr = pred.clone().requires_grad_(True)
        vr = O.masked_mse(pr, tgt, m, 32)
        good = abs(float(v.detach()) - float(vr.detach())) <= 1e-5 * max(1.0, abs(float(vr.detach())))
        print(f"[{'PASS' if good else 'FAIL'}] masked_mse {name}: {float(v.detach()):.6f} vs {float(vr.detach()):.6f}", flush=True)
        ok &= good
        if vr.requires_grad:
            v.backward()
            vr.backward()
            # a sample with an empty mask gets a NaN gradient from torch autograd (0/0 inside nanmean's
            # input); the fused kernel gives it a zero gradient.  Compare the non-empty samples.
            live = slice(1, None) if name == "one_empty_sample" else slice(None)
            ok &= report("   mse grad", p.grad[live].flatten(0, 2), pr.grad[live].flatten(0, 2), tol=1e-4)
            if name == "one_empty_sample":
                ok &= bool((p.grad[0] == 0).all())
        for crit, eps in ((ce, 0.0), (ce_ls, 0.1)):
            lg = logits.to(dev).requires_grad_(True)
            v = crit(lg, labels.to(dev), mask=md)
            lr = logits.clone().requires_grad_(True)
            vr = O.masked_ce(lr, labels, m, 8, eps)
            good = abs(float(v.detach()) - float(vr.detach())) <= 1e-5 * max(1.0, abs(float(vr.detach())))
            print(f"[{'PASS' if good else 'FAIL'}] masked_ce eps={eps} {name}: {float(v.detach()):.6f} vs {float(vr.detach()):.6f}", flush=True)
            ok &= good
            if vr.requires_grad:
                v.backward()
                vr.backward()
                live = slice(1, None) if name == "one_empty_sample" else slice(None)
                ok &= report("   ce grad", lg.grad[live].flatten(0, 2), lr.grad[live].flatten(0, 2), tol=1e-4)
                if name == "one_empty_sample":
                    ok &= bool((lg.grad[0] == 0).all())
    # ignore_index (-100) pixels: zero loss / zero gradient, still in the denominator -- against the oracle AND
    # against F.cross_entropy itself (the call the reference makes, criterion.py:33)
    lab_ig = labels.clone()
    lab_ig[torch.rand(lab_ig.shape, generator=g) < 0.2] = -100
    m = masks["random"]
    for crit, eps in ((ce, 0.0), (ce_ls, 0.1)):
        lg = logits.to(dev).requires_grad_(True)
        v = crit(lg, lab_ig.to(dev), mask=m.to(dev))
        lr = logits.clone().requires_grad_(True)
        vr = O.masked_ce(lr, lab_ig, m, 8, eps)
        per_px = torch.nn.functional.cross_entropy(logits, lab_ig, reduction="none", label_smoothing=eps)
        mu = m.view(B, 16, 16).repeat_interleave(8, 1).repeat_interleave(8, 2).float()
        vt = ((per_px * mu).flatten(1).sum(1) / mu.flatten(1).sum(1)).nanmean()
        good = abs(float(v.detach()) - float(vr.detach())) <= 1e-5 and abs(float(vr.detach()) - float(vt)) <= 1e-5
        print(f"[{'PASS' if good else 'FAIL'}] masked_ce eps={eps} ignore_index: {float(v.detach()):.6f} vs "
              f"{float(vr.detach()):.6f} (torch {float(vt):.6f})", flush=True)
        ok &= good
        v.backward()
        vr.backward()
        ok &= report("   ce grad (ignore_index)", lg.grad.flatten(0, 2), lr.grad.flatten(0, 2), tol=1e-4)
        ok &= bool((lg.grad.permute(0, 2, 3, 1)[lab_ig.to(dev) == -100] == 0).all())
    # a label outside [0, C) must not pass silently (torch raises a device assert): loss = +inf
    bad = labels.clone()
    bad[0, 0, 0] = 99
    v = ce(logits.to(dev), bad.to(dev), mask=torch.ones(B, 256, dtype=torch.long, device=dev))
    good = bool(torch.isinf(v))
    print(f"[{'PASS' if good else 'FAIL'}] masked_ce out-of-range label -> {float(v)}", flush=True)
    ok &= good
    # a sample whose loss is NaN is skipped by the masked form's nanmean (criterion.py:49/:107)
    pn = pred.clone()
    pn[1, 0, 5, 5] = float("nan")
    mo = torch.ones(B, 256, dtype=torch.long)
    v = mse(pn.to(dev), tgt.to(dev), mask=mo.to(dev))
    vr = O.masked_mse(pn, tgt, mo, 32)
    good = abs(float(v) - float(vr)) <= 1e-5 and not bool(torch.isnan(v))
    print(f"[{'PASS' if good else 'FAIL'}] masked_mse NaN sample skipped: {float(v):.6f} vs {float(vr):.6f}", flush=True)
    ok &= good
    v = mse(pn.to(dev), tgt.to(dev), mask=None)
    good = bool(torch.isnan(v))
    print(f"[{'PASS' if good else 'FAIL'}] masked_mse without mask propagates NaN: {float(v)}", flush=True)
    ok &= good
    return ok


def group_pretrain_tiny():
    from pretrain_case import run_pretrain_parity
    try:
        print(run_pretrain_parity(dev, "tiny", 3, verbose=True))
        return True
    except AssertionError as e:
        print("FAIL", str(e)[:600])
        return False


def group_pretrain_base():
    from pretrain_case import run_pretrain_parity
    try:
        print(run_pretrain_parity(dev, "base", 2, verbose=False))
        return True
    except AssertionError as e:
        print("FAIL", str(e)[:600])
        return False


if __name__ == "__main__":
    grp = sys.argv[1]
    t0 = time.time()
    print(f"=== group {grp} on {torch.cuda.get_device_name(0)} ===", flush=True)
    ok = globals()[f"group_{grp}"]()
    torch.cuda.synchronize()
    print(f"=== group {grp}: {'ALL PASS' if ok else 'FAILURES'} ({time.time() - t0:.1f}s) ===", flush=True)
    sys.exit(0 if ok else 1)
