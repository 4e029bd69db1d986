"""The CPU oracle (oracle/mirage_oracle.py) against the golden vectors produced by the unmodified
reference (oracle/make_golden.py).  Runs without a GPU; this is what pins the oracle."""
import torch

from helpers import GOLDEN, synth_images, synth_state_dict
from oracle import mirage_oracle as O
from pretrain_case import MODS, SIZES, build_pretrain_model, oracle_step


def _check_sub(got: torch.Tensor, gold: dict, rtol: float, atol: float = 0.0):
    assert tuple(got.shape) == tuple(gold["shape"])
    f = got.detach().float().reshape(-1, got.shape[-1])
    rows = f[:: gold["step"]]
    scale = gold["rows"].abs().max().item() + 1e-12
    assert (rows - gold["rows"]).abs().max().item() <= rtol * scale + atol
    assert abs(f.double().sum().item() - gold["sum"]) <= rtol * (gold["abs_sum"] + 1e-9)
    assert abs(f.double().abs().sum().item() - gold["abs_sum"]) <= rtol * (gold["abs_sum"] + 1e-9)


def _encoder_sd(size):
    from mirage_b200.mirage_hf import MIRAGEWrapper
    m = MIRAGEWrapper(size=size)
    sd = m.model.state_dict()
    sd.update(synth_state_dict({k: v.shape for k, v in sd.items()}, 0))
    return sd


def test_sincos_posemb_matches_host_module():
    from mirage_b200.utils import build_2d_sincos_posemb
    for (h, w, d) in [(16, 16, 768), (16, 16, 256), (8, 12, 64)]:
        assert torch.equal(O.sincos_posemb_2d(h, w, d), build_2d_sincos_posemb(h, w, d))


def test_encoder_base_golden():
    g = torch.load(GOLDEN / "encoder_base.pt")
    x = synth_images(g["batch"], ["bscan", "slo"], seed=g["input_seed"])
    with torch.no_grad():
        out = O.light_forward(x, _encoder_sd("base"), 12, 12)
    _check_sub(out, g["out"], 2e-5)


def test_encoder_large_golden():
    g = torch.load(GOLDEN / "encoder_large.pt")
    x = synth_images(g["batch"], ["bscan", "slo"], seed=g["input_seed"])
    with torch.no_grad():
        out = O.light_forward(x, _encoder_sd("large"), 24, 16)
    _check_sub(out, g["out"], 2e-5)


def test_random_masks_bit_exact():
    g = torch.load(GOLDEN / "masks.pt")
    for case in g["cases"]:
        torch.manual_seed(case["seed"])
        tm, keep, restore = O.random_masks(g["counts"], case["B"], case["n_vis"], alphas=case["alphas"])
        assert torch.equal(keep, case["ids_keep"].long())
        assert torch.equal(restore, case["ids_restore"].long())
        for d, m in zip(g["domains"], tm):
            assert torch.equal(m, case["task_masks"][d].long())
        # invariants the reference guarantees
        assert int(sum(int((m == 0).sum()) for m in tm)) == case["B"] * case["n_vis"]
        assert torch.equal(torch.gather(restore, 1, keep),
                           torch.arange(case["n_vis"]).expand(case["B"], -1))


def test_criterion_golden():
    g = torch.load(GOLDEN / "criterion.pt")
    gen = torch.Generator().manual_seed(g["seed"])
    B = g["B"]
    pred = torch.randn(B, 1, 512, 512, generator=gen)
    tgt = torch.rand(B, 1, 512, 512, generator=gen)
    logits = torch.randn(B, 13, 128, 128, generator=gen)
    labels = torch.randint(0, 13, (B, 128, 128), generator=gen)
    for k in g["mse"]:
        m = None if k == "no_mask" else g["masks"][k].long()
        assert abs(float(O.masked_mse(pred, tgt, m, 32)) - g["mse"][k]) <= 1e-5 * max(1.0, abs(g["mse"][k]))
        assert abs(float(O.masked_ce(logits, labels, m, 8)) - g["ce"][k]) <= 1e-5 * max(1.0, abs(g["ce"][k]))
        assert abs(float(O.masked_ce(logits, labels, m, 8, 0.1)) - g["ce_ls"][k]) <= 1e-5 * max(1.0, abs(g["ce_ls"][k]))


def _pretrain_golden(tag):
    g = torch.load(GOLDEN / f"pretrain_{tag}.pt")
    assert (g["dim"], g["depth"], g["heads"]) == SIZES[tag]
    model, _ = build_pretrain_model(tag)
    sd = model.state_dict()
    sd.update(synth_state_dict({k: v.shape for k, v in sd.items()}, g["weights_seed"]))
    x = synth_images(g["batch"], MODS, seed=g["input_seed"])
    masks = ({k: v.long() for k, v in g["task_masks"].items()}, g["ids_keep"].long(), g["ids_restore"].long())
    preds, losses, grads = oracle_step(sd, x, masks, tag)
    for d in MODS:
        assert abs(losses[d] - g["losses"][d]) <= 2e-5 * abs(g["losses"][d]), (d, losses[d], g["losses"][d])
        _check_sub(preds[d], g["preds"][d], 5e-5)
    assert set(grads) == set(g["grad_norm"])
    for k, n in g["grad_norm"].items():
        assert abs(grads[k].norm().item() - n) <= 1e-3 * n + 1e-7, (k, grads[k].norm().item(), n)
    for k, t in g["grad_small"].items():
        assert (grads[k] - t).abs().max().item() <= 1e-3 * (t.abs().max().item() + 1e-9) + 1e-8, k
    for k, sub in g["grad_big"].items():
        gk = grads[k]
        _check_sub(gk.reshape(-1, gk.shape[-1]) if gk.dim() > 1 else gk.reshape(1, -1),
                   {**sub, "shape": tuple((gk.reshape(-1, gk.shape[-1]) if gk.dim() > 1 else gk.reshape(1, -1)).shape)},
                   2e-3, atol=1e-8)  # atol: gradients that are analytically zero (key bias) are pure noise


def test_pretrain_tiny_golden():
    _pretrain_golden("tiny")


def test_pretrain_base_golden():
    _pretrain_golden("base")


def test_pretrain_large_golden():
    """BASELINE configs[3] at its real model size (ViT-L + three decoders, batch 2)."""
    _pretrain_golden("large")


def test_cls_large_golden():
    """BASELINE configs[4] at its real model size: miragecls_factory['global'] on the ViT-L encoder."""
    from cls_case import build_cls_model, oracle_cls_logits
    g = torch.load(GOLDEN / "cls_large.pt")
    assert g["size"] == "large"
    x = synth_images(2, ["bscan"], seed=g["input_seed"])["bscan"]
    m, sd = build_cls_model("global", g["weights_seed"], size="large")
    assert sum(p.numel() for p in m.parameters()) == g["out"]["global"]["n_params"]
    with torch.no_grad():
        logits = oracle_cls_logits(x, sd, "global", size="large")
    ref = g["out"]["global"]["logits"]
    assert torch.allclose(logits, ref, rtol=2e-3, atol=2e-3), (logits - ref).abs().max()


def test_cls_heads_golden():
    """Oracle of the classification tail (mirage_wrapper.py:187-244) against logits recorded from the
    reference's miragecls_factory classes (ViT-B encoder, B=2)."""
    import torch
    from cls_case import build_cls_model, oracle_cls_logits
    from helpers import GOLDEN, synth_images
    g = torch.load(GOLDEN / "cls.pt")
    x = synth_images(2, ["bscan"], seed=g["input_seed"])["bscan"]
    for pool in ("global", "cls", "token_mix"):
        m, sd = build_cls_model(pool, g["weights_seed"])
        assert sum(p.numel() for p in m.parameters()) == g["out"][pool]["n_params"]
        with torch.no_grad():
            logits = oracle_cls_logits(x, sd, pool)
        ref = g["out"][pool]["logits"]
        assert torch.allclose(logits, ref, rtol=2e-3, atol=2e-3), (pool, (logits - ref).abs().max())
        loss = torch.nn.functional.cross_entropy(logits, torch.tensor([1, 3])).item()
        assert abs(loss - g["out"][pool]["loss"]) <= 2e-3 * abs(g["out"][pool]["loss"])


def test_seg_heads_golden():
    """Oracle of the segmentation heads (mirage/output_adapters.py:437-575) against predictions recorded from the
    reference's MIRAGELight + LinearSegAdapter / ConvNeXtAdapter; the product modules expose the same keys."""
    from seg_case import build_seg_model, oracle_seg
    g = torch.load(GOLDEN / "seg.pt")
    x = synth_images(g["batch"], ["bscan"], seed=g["input_seed"])
    for kind in ("linear", "convnext"):
        m = build_seg_model(kind)
        sd = m.state_dict()
        assert sorted(sd.keys()) == g["out"][kind]["keys"]
        assert sum(p.numel() for p in m.parameters()) == g["out"][kind]["n_params"]
        sd.update(synth_state_dict({k: v.shape for k, v in sd.items()}, g["weights_seed"]))
        with torch.no_grad():
            pred = oracle_seg(x, sd, kind)
        _check_sub(pred, g["out"][kind]["pred"], 5e-5)


def test_checkpoint_schema_round_trip(tmp_path):
    """mutils/checkpoint.py:9-72: save_model writes the reference's dictionary, auto_load_model resumes from the
    newest numeric checkpoint and restores model / optimizer / epoch; the cls wrappers rebuild the model from the
    pickled args of such a file (mirage_wrapper.py:59-62)."""
    import argparse

    from mirage_b200.checkpoint import auto_load_model, latest_checkpoint, save_model
    from mirage_b200.optim import NativeScalerWithGradNormCount
    from seg_case import build_seg_model
    m = build_seg_model("linear")
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    for p in m.parameters():
        if p.requires_grad:
            p.grad = torch.randn_like(p)
    opt.step()
    args = argparse.Namespace(output_dir=str(tmp_path), auto_resume=True, resume="", model="miragelight_tiny")
    scaler = NativeScalerWithGradNormCount()
    for epoch in (3, 11):
        path = save_model(args, epoch, m, opt, scaler)
    assert latest_checkpoint(tmp_path).endswith("checkpoint-11.pth")
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) == {"model", "optimizer", "epoch", "scaler", "args"} and ck["epoch"] == 11
    assert ck["args"].model == "miragelight_tiny" and ck["scaler"] == {"scale": 1.0}
    m2 = build_seg_model("linear")
    opt2 = torch.optim.AdamW(m2.parameters(), lr=1e-3)
    args2 = argparse.Namespace(output_dir=str(tmp_path), auto_resume=True, resume="")
    auto_load_model(args2, m2, opt2, scaler)
    assert args2.start_epoch == 12 and args2.resume.endswith("checkpoint-11.pth")
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert s1.keys() == s2.keys() and all(torch.equal(s1[i]["exp_avg"], s2[i]["exp_avg"]) for i in s1)


def test_semseg_interp_golden():
    """SemSegInputAdapter(interpolate_class_emb=True): oracle restatement vs the reference's recorded tokens and
    parameter gradients."""
    fx = torch.load(GOLDEN / "semseg_interp.pt")
    sd = {k: v.clone().requires_grad_(k != "pos_emb") for k, v in fx["state_dict"].items()}
    tok = O.semseg_embed_interp(fx["labels"], sd["class_emb.weight"], sd["proj.1.weight"], sd["proj.1.bias"],
                                sd["pos_emb"], (8, 8))
    assert (tok - fx["tokens"]).abs().max().item() <= 1e-5 * fx["tokens"].abs().max().item()
    (tok * fx["cotangent"]).sum().backward()
    for k, g in fx["grads"].items():
        assert (sd[k].grad - g).abs().max().item() <= 1e-4 * max(1e-6, g.abs().max().item()), k
