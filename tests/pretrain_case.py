"""MultiMAE pretraining case shared by the GPU parity tests, __graft_entry__.smoke() and bench.py:
builds the model (mirage_b200 modules), and runs forward + masked losses + backward against the CPU
oracle on identical weights, inputs and recorded masks."""
from __future__ import annotations

import argparse

import torch

from helpers import load_synth, parity, synth_images

MODS = ["bscan", "slo", "bscanlayermap"]
SIZES = {"tiny": (128, 2, 2), "base": (768, 12, 12), "large": (1024, 24, 16)}


def pretrain_args(mods=MODS):
    a = argparse.Namespace()
    a.in_domains = list(mods)
    a.out_domains = list(mods)
    a.patch_size = {d: ((8, 8) if d == "bscanlayermap" else (32, 32)) for d in mods}
    a.input_size = {d: ((128, 128) if d == "bscanlayermap" else (512, 512)) for d in mods}
    a.grid_sizes = {d: [16, 16] for d in mods}
    return a


def build_pretrain_model(size="base", mods=MODS):
    from mirage_b200.input_adapters import PatchedInputAdapter, SemSegInputAdapter
    from mirage_b200.model import MIRAGEModel
    from mirage_b200.output_adapters import SpatialOutputAdapter
    dim, depth, heads = SIZES[size]
    args = pretrain_args(mods)
    ins, outs = {}, {}
    for d in mods:
        if d == "bscanlayermap":
            ins[d] = SemSegInputAdapter(num_classes=13, stride_level=1, patch_size_full=(8, 8),
                                        image_size=(128, 128), dim_class_emb=64, interpolate_class_emb=False)
            ch = 13
        else:
            ins[d] = PatchedInputAdapter(num_channels=1, stride_level=1, patch_size_full=(32, 32),
                                         image_size=(512, 512))
            ch = 1
        outs[d] = SpatialOutputAdapter(num_channels=ch, stride_level=1, patch_size_full=tuple(args.patch_size[d]),
                                       dim_tokens=256, depth=2, num_heads=8, use_task_queries=True, task=d,
                                       context_tasks=list(mods), use_xattn=True, image_size=args.input_size[d])
    model = MIRAGEModel(args, input_adapters=ins, output_adapters=outs, num_global_tokens=1,
                        dim_tokens=dim, depth=depth, num_heads=heads, drop_path_rate=0.0)
    return model, args


def build_criteria(mods=MODS):
    from mirage_b200.criterion import MaskedCrossEntropyLoss, MaskedMSELoss
    return {d: (MaskedCrossEntropyLoss(patch_size=(8, 8), stride=1) if d == "bscanlayermap"
                else MaskedMSELoss(patch_size=(32, 32), stride=1)) for d in mods}


def oracle_step(sd, x, masks, size, mods=MODS):
    """Oracle forward + losses + autograd gradients (fp32, CPU).  Returns (preds, losses, grads)."""
    from oracle import mirage_oracle as O
    dim, depth, heads = SIZES[size]
    leaf = {k: v.clone().requires_grad_(not k.endswith("pos_emb")) for k, v in sd.items()}
    task_masks, ids_keep, ids_restore = masks
    preds, _ = O.pretrain_forward(x, leaf, depth, heads, (None, ids_keep, ids_restore), mods)
    total, losses = O.pretrain_loss(preds, x, task_masks)
    total.backward()
    grads = {k: v.grad for k, v in leaf.items() if v.grad is not None}
    return preds, {d: float(v.detach()) for d, v in losses.items()}, grads


def b200_step(model, crits, x_dev, masks_dev, mods=MODS):
    """mirage_b200 forward + losses + backward with injected masks.  Returns (preds, losses, grads)."""
    task_masks, ids_keep, ids_restore = masks_dev
    model.generate_random_masks = lambda *a, **k: (task_masks, ids_keep, ids_restore)
    model.zero_grad(set_to_none=True)
    preds, masks = model(x_dev, num_encoded_tokens=ids_keep.shape[1], alphas=1.0, sample_tasks_uniformly=False)
    losses = {d: crits[d](preds[d].float(), x_dev[d], mask=masks[d]) for d in mods}
    total = sum(losses.values())
    total.backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    return preds, {d: float(v.detach()) for d, v in losses.items()}, grads


def sample_masks(model, batch, n_vis, seed, device="cpu"):
    torch.manual_seed(seed)
    toks = {d: torch.empty(batch, 256, 0, device=device) for d in MODS}
    return model.generate_random_masks(toks, n_vis, alphas=1.0)


def run_pretrain_parity(dev, size="tiny", batch=2, verbose=False):
    """Returns a dict of worst-case parity metrics; raises AssertionError outside the bf16 tolerance."""
    model, _ = build_pretrain_model(size)
    sd = load_synth(model, seed=3)
    model = model.to(dev).train()
    crits = build_criteria()
    x = synth_images(batch, MODS, seed=77)
    tm, keep, restore = sample_masks(model, batch, 98, seed=11)
    masks = (tm, keep, restore)
    p_ref, l_ref, g_ref = oracle_step(sd, x, masks, size)
    x_dev = {k: v.to(dev) for k, v in x.items()}
    masks_dev = ({k: v.to(dev) for k, v in tm.items()}, keep.to(dev), restore.to(dev))
    p, l, g = b200_step(model, crits, x_dev, masks_dev)
    torch.cuda.synchronize()
    out = {"loss_rel": 0.0, "pred_max_rel": 0.0, "pred_min_cos": 1.0, "grad_max_rel_fro": 0.0, "grad_min_cos": 1.0}
    for d in MODS:
        m = parity(p[d].reshape(-1, p[d].shape[-1]), p_ref[d].reshape(-1, p_ref[d].shape[-1]))  # rows = image rows
        out["pred_max_rel"] = max(out["pred_max_rel"], m["max_rel"])
        out["pred_min_cos"] = min(out["pred_min_cos"], m["min_cos"])
        out["loss_rel"] = max(out["loss_rel"], abs(l[d] - l_ref[d]) / abs(l_ref[d]))
        if verbose:
            print(f"  pred {d}: {m}  loss {l[d]:.5f} vs {l_ref[d]:.5f}")
    worst = None
    assert set(g) == set(g_ref), (sorted(set(g_ref) - set(g))[:5], sorted(set(g) - set(g_ref))[:5])
    biggest = max(v.norm().item() for v in g_ref.values())
    for k in g_ref:
        a, b = g[k].detach().float().cpu().flatten(), g_ref[k].flatten()
        if b.norm().item() < 1e-5 * biggest:
            # analytically-zero gradients (attention key bias: softmax is shift invariant) are
            # rounding noise in the fp32 oracle too; only require them to stay negligible
            assert a.norm().item() < 1e-3 * biggest, (k, a.norm().item(), biggest)
            continue
        rel = ((a - b).norm() / (b.norm() + 1e-12)).item()
        cos = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
        if verbose:
            print(f"  grad {k}: rel_fro={rel:.4f} cos={cos:.5f} |ref|={b.norm():.3e}")
        if rel > out["grad_max_rel_fro"]:
            out["grad_max_rel_fro"], worst = rel, k
        out["grad_min_cos"] = min(out["grad_min_cos"], cos)
    out["worst_grad"] = worst
    assert out["loss_rel"] <= 2e-2, out
    assert out["pred_max_rel"] <= 2e-2 and out["pred_min_cos"] >= 0.999, out
    assert out["grad_max_rel_fro"] <= 5e-2 and out["grad_min_cos"] >= 0.999, out
    return out
