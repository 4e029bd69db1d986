"""Generate tests/golden/*.pt from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference (read-only mount)

The reference cannot travel to the GPU box, so its outputs on seeded synthetic inputs / weights are
committed as small fixtures.  Weights are NOT stored: tests/helpers.synth_state_dict regenerates
them bit-identically from the seed (CPU mt19937).  Large outputs are stored as strided row subsets
plus whole-tensor checksums.

Fixtures
  encoder_{base,large}.pt  hf/mirage_hf.py MIRAGEWrapper forward, bscan+slo 512x512, B=1
  masks.pt                 MIRAGEModel.generate_random_masks under fixed seeds (bit-exact target)
  pretrain_{tiny,base,large}.pt  MIRAGEModel (+3 SpatialOutputAdapters) forward, masked losses, gradients
  criterion.pt             MaskedMSELoss / MaskedCrossEntropyLoss incl. empty- and partial-mask cases
  cls.pt                   miragecls_factory['global'|'cls'|'token_mix'] logits + grads (ViT-B encoder)
  cls_large.pt             miragecls_factory['global'] logits + grads, ViT-L encoder (BASELINE configs[4])
  seg.pt                   MIRAGELight + LinearSegAdapter / ConvNeXtAdapter predictions (seg-tuning caller)
"""
from __future__ import annotations

import argparse
import contextlib
import io
import sys
import types
from functools import partial
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import GOLDEN, synth_images, synth_state_dict  # noqa: E402


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def import_reference():
    assert REF.exists(), "reference checkout not mounted"
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(REF / "hf"))
    # mirage_wrapper.py imports skimage (absent here) only for its file-loading helper
    for name in ("skimage", "skimage.io", "skimage.transform"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["skimage.transform"].resize = lambda *a, **k: None
    sys.modules["skimage"].io = sys.modules["skimage.io"]
    with _quiet():
        import mirage_hf as ref_hf
        import mirage.criterion as ref_crit
        import mirage.input_adapters as ref_in
        import mirage.model as ref_model
        import mirage.output_adapters as ref_out
        import mirage_wrapper as ref_wrap
    return ref_hf, ref_model, ref_in, ref_out, ref_crit, ref_wrap


def subsample(t: torch.Tensor, step: int):
    """rows 0, step, 2*step, ... of the flattened [-1, last] view + checksums of the whole tensor."""
    f = t.detach().float().reshape(-1, t.shape[-1])
    return {"shape": tuple(t.shape), "step": step, "rows": f[::step].clone(),
            "sum": f.double().sum().item(), "abs_sum": f.double().abs().sum().item()}


def load_synth_into(model, seed):
    sd = model.state_dict()
    sd.update({k: v.to(sd[k].dtype) for k, v in synth_state_dict({k: v.shape for k, v in sd.items()}, seed).items()})
    model.load_state_dict(sd)
    return sd


def pretrain_args(mods):
    a = argparse.Namespace()
    a.in_domains = list(mods)
    a.out_domains = list(mods)
    a.patch_size = {d: ((8, 8) if d == "bscanlayermap" else (32, 32)) for d in mods}
    a.input_size = {d: ((128, 128) if d == "bscanlayermap" else (512, 512)) for d in mods}
    a.grid_sizes = {d: [16, 16] for d in mods}
    return a


def build_ref_pretrain(ref_model, ref_in, ref_out, mods, dim, depth, heads):
    args = pretrain_args(mods)
    with _quiet():
        ins, outs = {}, {}
        for d in mods:
            if d == "bscanlayermap":
                ins[d] = ref_in.SemSegInputAdapter(num_classes=13, stride_level=1, patch_size_full=(8, 8),
                                                   image_size=(128, 128), dim_class_emb=64,
                                                   interpolate_class_emb=False)
                ch = 13
            else:
                ins[d] = ref_in.PatchedInputAdapter(num_channels=1, stride_level=1, patch_size_full=(32, 32),
                                                    image_size=(512, 512))
                ch = 1
            outs[d] = ref_out.SpatialOutputAdapter(
                num_channels=ch, stride_level=1, patch_size_full=tuple(args.patch_size[d]), dim_tokens=256,
                depth=2, num_heads=8, use_task_queries=True, task=d, context_tasks=list(mods),
                use_xattn=True, image_size=args.input_size[d])
        model = ref_model.MIRAGEModel(args, input_adapters=ins, output_adapters=outs, num_global_tokens=1,
                                      dim_tokens=dim, depth=depth, num_heads=heads, drop_path_rate=0.0)
    return model, args


def gen_encoder(ref_hf, size):
    with _quiet():
        m = ref_hf.MIRAGEWrapper(size=size).eval()
    load_synth_into(m.model, 0)
    x = synth_images(1, ["bscan", "slo"], seed=1234)
    with torch.no_grad():
        out = m(dict(x))
    torch.save({"weights_seed": 0, "input_seed": 1234, "batch": 1, "out": subsample(out, 8)},
               GOLDEN / f"encoder_{size}.pt")
    print("encoder", size, tuple(out.shape))


def gen_masks(ref_model, ref_in, ref_out):
    mods = ["bscan", "slo", "bscanlayermap"]
    model, _ = build_ref_pretrain(ref_model, ref_in, ref_out, mods, 128, 1, 2)
    cases = []
    for (seed, B, n_vis, alphas) in [(0, 4, 98, 1.0), (1, 7, 98, 1.0), (2, 3, 128, 0.5), (3, 2, 768, 1.0), (4, 5, 1, 1.0)]:
        torch.manual_seed(seed)
        toks = {d: torch.empty(B, 256, 0) for d in mods}
        with _quiet():
            tm, keep, restore = model.generate_random_masks(toks, n_vis, alphas=alphas)
        cases.append({"seed": seed, "B": B, "n_vis": n_vis, "alphas": alphas,
                      "task_masks": {k: v.to(torch.int8) for k, v in tm.items()},
                      "ids_keep": keep.to(torch.int16), "ids_restore": restore.to(torch.int16)})
    torch.save({"cases": cases, "counts": [256, 256, 256], "domains": mods}, GOLDEN / "masks.pt")
    print("masks", len(cases))


def gen_pretrain(ref_model, ref_in, ref_out, ref_crit, tag, dim, depth, heads, B, big_step):
    mods = ["bscan", "slo", "bscanlayermap"]
    model, args = build_ref_pretrain(ref_model, ref_in, ref_out, mods, dim, depth, heads)
    load_synth_into(model, 3)
    model.train()
    x = synth_images(B, mods, seed=77)
    torch.manual_seed(11)
    toks = {d: torch.empty(B, 256, 0) for d in mods}
    with _quiet():
        task_masks, ids_keep, ids_restore = model.generate_random_masks(toks, 98, alphas=1.0)
    # inject the recorded masks (the reference's own task_masks= argument is B=1-only)
    model.generate_random_masks = lambda *a, **k: (task_masks, ids_keep, ids_restore)
    with _quiet():
        preds, masks = model(dict(x), num_encoded_tokens=98, alphas=1.0, sample_tasks_uniformly=False)
    losses = {}
    for d in mods:
        fn = (ref_crit.MaskedCrossEntropyLoss(patch_size=(8, 8), stride=1) if d == "bscanlayermap"
              else ref_crit.MaskedMSELoss(patch_size=(32, 32), stride=1))
        losses[d] = fn(preds[d].float(), x[d], mask=masks[d])
    total = sum(losses.values())
    total.backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    small = {k: g.clone() for k, g in grads.items() if g.numel() <= 4096}
    big = {k: subsample(g.reshape(-1, g.shape[-1]) if g.dim() > 1 else g.reshape(1, -1), big_step)
           for k, g in grads.items() if g.numel() > 4096}
    torch.save({
        "weights_seed": 3, "input_seed": 77, "batch": B, "dim": dim, "depth": depth, "heads": heads,
        "task_masks": {k: v.to(torch.int8) for k, v in task_masks.items()},
        "ids_keep": ids_keep.to(torch.int16), "ids_restore": ids_restore.to(torch.int16),
        "preds": {d: subsample(preds[d], 64) for d in mods},
        "losses": {d: v.item() for d, v in losses.items()}, "loss": total.item(),
        "grad_norm": {k: g.norm().item() for k, g in grads.items()},
        "grad_small": small, "grad_big": big,
        "no_grad_params": [k for k, p in model.named_parameters() if p.grad is None],
    }, GOLDEN / f"pretrain_{tag}.pt")
    print("pretrain", tag, {d: round(v.item(), 5) for d, v in losses.items()})


def gen_criterion(ref_crit):
    g = torch.Generator().manual_seed(5)
    B = 3
    pred = torch.randn(B, 1, 512, 512, generator=g)
    tgt = torch.rand(B, 1, 512, 512, generator=g)
    logits = torch.randn(B, 13, 128, 128, generator=g)
    labels = torch.randint(0, 13, (B, 128, 128), generator=g)
    masks = {
        "random": (torch.rand(B, 256, generator=g) > 0.3).long(),
        "all_masked": torch.ones(B, 256, dtype=torch.long),
        "none_masked": torch.zeros(B, 256, dtype=torch.long),
        "one_empty_sample": torch.cat([torch.zeros(1, 256, dtype=torch.long),
                                       (torch.rand(B - 1, 256, generator=g) > 0.5).long()]),
    }
    mse = ref_crit.MaskedMSELoss(patch_size=(32, 32), stride=1)
    ce = ref_crit.MaskedCrossEntropyLoss(patch_size=(8, 8), stride=1)
    ce_ls = ref_crit.MaskedCrossEntropyLoss(patch_size=(8, 8), stride=1, label_smoothing=0.1)
    out = {"seed": 5, "B": B, "masks": {k: v.to(torch.int8) for k, v in masks.items()}, "mse": {}, "ce": {},
           "ce_ls": {}, "mse_grad": {}, "ce_grad": {}}
    for k, m in list(masks.items()) + [("no_mask", None)]:
        p = pred.clone().requires_grad_(True)
        v = mse(p, tgt, mask=m)
        out["mse"][k] = float(v.detach())
        if v.requires_grad:
            v.backward()
            out["mse_grad"][k] = subsample(p.grad, 4096)
        lg = logits.clone().requires_grad_(True)
        v = ce(lg, labels, mask=m)
        out["ce"][k] = float(v.detach())
        if v.requires_grad:
            v.backward()
            out["ce_grad"][k] = subsample(lg.grad, 1024)
        out["ce_ls"][k] = float(ce_ls(logits, labels, mask=m))
    torch.save(out, GOLDEN / "criterion.pt")
    print("criterion", out["mse"], out["ce"])


def gen_cls(ref_wrap, ref_model, ref_in, size="base", pools=("global", "cls", "token_mix"), fname="cls.pt"):
    """Classification heads on the ViT-B (cls.pt) / ViT-L (cls_large.pt, BASELINE configs[4]) encoder:
    the wrapper classes need a checkpoint file, so one is synthesised in /tmp with the recipe of
    SURVEY.md 8(c)."""
    import tempfile
    args = argparse.Namespace(model=f"miragepre_{size}", out_domains=[], decoder_dim=256, decoder_depth=2,
                              decoder_num_heads=8, decoder_use_task_queries=True, decoder_use_xattn=True,
                              num_global_tokens=1, drop_path=0.0, grid_sizes={"bscan": [16, 16]})
    out = {}
    for pool in pools:
        cls_t = ref_wrap.miragecls_factory[pool]
        with _quiet():
            helper = cls_t.__new__(cls_t)
            torch.nn.Module.__init__(helper)
            a2 = argparse.Namespace(**vars(args))
            a2.in_domains = ["bscan"]
            a2.patch_size = {"bscan": (32, 32)}
            a2.input_size = {"bscan": (512, 512)}
            helper.args = a2
            enc = helper.get_model()
        with tempfile.NamedTemporaryFile(suffix=".pth") as f:
            torch.save({"model": enc.state_dict(), "args": args}, f.name)
            with _quiet():
                m = cls_t(num_classes=5, input_size=512, patch_size=32, modalities="bscan", weights=f.name,
                          device="cpu")
        sd = load_synth_into(m, 21)
        m.train()
        x = synth_images(2, ["bscan"], seed=9)["bscan"]
        torch.manual_seed(13)   # the encoder draws a random token permutation
        with _quiet():
            logits = m(x)
        tgt = torch.tensor([1, 3])
        loss = torch.nn.functional.cross_entropy(logits, tgt)
        loss.backward()
        out[pool] = {"logits": logits.detach().clone(), "loss": loss.item(),
                     "grad_norm": {k: p.grad.norm().item() for k, p in m.named_parameters() if p.grad is not None},
                     "head_grad": m.head.weight.grad.clone(), "n_params": sum(p.numel() for p in m.parameters())}
        print("cls", pool, logits.detach().flatten()[:3].tolist(), loss.item())
    torch.save({"weights_seed": 21, "input_seed": 9, "mask_seed": 13, "size": size, "out": out}, GOLDEN / fname)


def gen_seg(ref_model, ref_in, ref_out):
    """MIRAGELight (small encoder: dim 128, depth 2) with the two segmentation heads of the seg-tuning caller
    (LinearSegAdapter, ConvNeXtAdapter; mirage/output_adapters.py:437-575), bscan 512x512, B = 2; also the
    encoder's return_all_layers list (model.py:545-553)."""
    a = argparse.Namespace()
    a.in_domains = ["bscan"]
    a.patch_size = {"bscan": (32, 32)}
    a.input_size = {"bscan": (512, 512)}
    a.grid_sizes = {"bscan": [16, 16]}
    out = {}
    for name, make in (("linear", lambda: ref_out.LinearSegAdapter(num_classes=13, main_tasks=("bscan",),
                                                                    patch_size=[32, 32], task="bscan",
                                                                    image_size=(512, 512))),
                       ("convnext", lambda: ref_out.ConvNeXtAdapter(num_classes=13, embed_dim=2048, preds_per_patch=16,
                                                                    main_tasks=("bscan",), patch_size=[32, 32],
                                                                    depth=2, task="bscan", image_size=(512, 512)))):
        with _quiet():
            ins = {"bscan": ref_in.PatchedInputAdapter(num_channels=1, stride_level=1, patch_size_full=(32, 32),
                                                       image_size=(512, 512))}
            m = ref_model.MIRAGELight(a, input_adapters=ins, output_adapters={"bscan": make()},
                                      num_global_tokens=1, dim_tokens=128, depth=2, num_heads=2,
                                      drop_path_rate=0.0).eval()
        load_synth_into(m, 31)
        x = synth_images(2, ["bscan"], seed=41)
        with torch.no_grad(), _quiet():
            pred = m(dict(x))["bscan"]
        out[name] = {"pred": subsample(pred, 128), "keys": sorted(m.state_dict().keys()),
                     "n_params": sum(p.numel() for p in m.parameters())}
        print("seg", name, tuple(pred.shape))
    torch.save({"weights_seed": 31, "input_seed": 41, "batch": 2, "out": out}, GOLDEN / "seg.pt")


def gen_semseg_interp(ref_in):
    """SemSegInputAdapter(interpolate_class_emb=True) (mirage/input_adapters.py:194-200, :226-238): bilinear
    down-sampling of the embedded class map + 1x1 projection; forward and the gradients of its parameters."""
    with _quiet():
        ad = ref_in.SemSegInputAdapter(num_classes=13, stride_level=1, patch_size_full=(8, 8), dim_tokens=128,
                                       image_size=(128, 128), dim_class_emb=64, interpolate_class_emb=True)
    sd = load_synth_into(ad, 17)
    g = torch.Generator().manual_seed(23)
    labels = torch.randint(0, 13, (2, 128, 128), generator=g)
    out = ad(labels)
    w = torch.randn(out.shape, generator=g)
    (out * w).sum().backward()
    grads = {k: p.grad.clone() for k, p in ad.named_parameters() if p.grad is not None}
    torch.save({"state_dict": {k: v.clone() for k, v in sd.items()}, "labels": labels, "tokens": out.detach(),
                "cotangent": w, "grads": grads}, GOLDEN / "semseg_interp.pt")
    print("semseg_interp", tuple(out.shape), sorted(grads))


if __name__ == "__main__":
    GOLDEN.mkdir(parents=True, exist_ok=True)
    which = set(sys.argv[1:]) or {"encoder", "masks", "pretrain", "pretrain_large", "criterion", "cls", "cls_large", "seg",
                                  "semseg_interp"}
    ref_hf, ref_model, ref_in, ref_out, ref_crit, ref_wrap = import_reference()
    torch.set_num_threads(8)
    if "encoder" in which:
        gen_encoder(ref_hf, "base")
        gen_encoder(ref_hf, "large")
    if "masks" in which:
        gen_masks(ref_model, ref_in, ref_out)
    if "criterion" in which:
        gen_criterion(ref_crit)
    if "pretrain" in which:
        gen_pretrain(ref_model, ref_in, ref_out, ref_crit, "tiny", 128, 2, 2, 3, 64)
        gen_pretrain(ref_model, ref_in, ref_out, ref_crit, "base", 768, 12, 12, 2, 1024)
    if "pretrain_large" in which:   # BASELINE configs[3] at its real model size (ViT-L), batch 2
        gen_pretrain(ref_model, ref_in, ref_out, ref_crit, "large", 1024, 24, 16, 2, 8192)
    if "cls" in which:
        gen_cls(ref_wrap, ref_model, ref_in)
    if "seg" in which:
        gen_seg(ref_model, ref_in, ref_out)
    if "semseg_interp" in which:
        gen_semseg_interp(ref_in)
    if "cls_large" in which:        # BASELINE configs[4] at its real model size (ViT-L), batch 2
        gen_cls(ref_wrap, ref_model, ref_in, size="large", pools=("global",), fname="cls_large.pt")
