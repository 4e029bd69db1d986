"""-m gpu: visible-token embedding (csrc/visible.cu, functional.embed_visible) -- the masked forward that embeds
only the kept patches -- against embed-everything-then-gather (the reference's order, mirage/model.py:352-391) on
the same weights, inputs and masks: index kernels exact, tokens to fp32 rounding, adapter gradients to bf16
summation order; and the whole pretraining step (both orders) against each other."""
import pytest
import torch

from helpers import load_synth, synth_images
from pretrain_case import MODS, b200_step, build_criteria, build_pretrain_model

pytestmark = pytest.mark.gpu


def _masks(B, n_keep, counts, seed, dev):
    g = torch.Generator().manual_seed(seed)
    n_all = sum(counts)
    ids_shuffle = torch.stack([torch.randperm(n_all, generator=g) for _ in range(B)])
    ids_restore = torch.argsort(ids_shuffle, dim=1)
    ids_keep = ids_shuffle[:, :n_keep].contiguous()
    mask_all = torch.ones(B, n_all, dtype=torch.long)
    mask_all.scatter_(1, ids_keep, 0)
    task_masks = dict(zip(MODS, torch.split(mask_all, counts, dim=1)))
    return ({k: v.to(dev) for k, v in task_masks.items()}, ids_keep.to(dev), ids_restore.to(dev))


def test_index_kernels_exact():
    from mirage_b200 import ops
    dev = torch.device("cuda:0")
    B, n_keep, n_glob, counts = 5, 37, 2, [256, 256, 256]
    _, ids_keep, _ = _masks(B, n_keep, counts, 3, dev)
    starts = [0, 256, 512]
    row_src, row_cls = ops.visible_rows(ids_keep, starts, counts, n_glob)
    ik = ids_keep.cpu()
    T = B * (n_keep + n_glob)
    want_src = torch.full((3, T), -1, dtype=torch.int32)
    want_cls = torch.empty(T, dtype=torch.int32)
    for b in range(B):
        for j in range(n_keep + n_glob):
            t = b * (n_keep + n_glob) + j
            if j >= n_keep:
                want_cls[t] = 3 + j - n_keep
                continue
            m = int(ik[b, j]) // 256
            want_cls[t] = m
            want_src[m, t] = b * 256 + int(ik[b, j]) % 256
    assert torch.equal(row_src.cpu(), want_src) and torch.equal(row_cls.cpu(), want_cls)

    g = torch.Generator(device="cuda").manual_seed(0)
    img = torch.rand(B, 1, 512, 512, device=dev, generator=g)
    a32, a16 = ops.gather_patches32(img, row_src[1], True, True)
    patches = img.reshape(B, 16, 32, 16, 32).permute(0, 1, 3, 2, 4).reshape(B * 256, 1024)
    src = row_src[1].long()
    want = torch.where((src >= 0)[:, None], patches[src.clamp(min=0)], torch.zeros((), device=dev))
    assert torch.equal(a32, want) and torch.equal(a16, want.to(torch.bfloat16))

    dy = torch.randn(T, 256, device=dev, generator=g)
    cs = ops.class_colsum(dy, row_cls, 5)
    for c in range(5):
        ref = dy[row_cls == c].double().sum(0)
        assert (cs[c].double() - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("size,B,n_keep", [("tiny", 6, 98), ("base", 3, 41)])
def test_masked_step_matches_embed_all_then_gather(size, B, n_keep):
    dev = torch.device("cuda:0")
    model, _ = build_pretrain_model(size)
    load_synth(model, seed=3)
    model = model.to(dev).train()
    crits = build_criteria()
    x = {k: v.to(dev) for k, v in synth_images(B, MODS, seed=11).items()}
    masks = _masks(B, n_keep, [256, 256, 256], 5, dev)

    # the embedding alone: same token rows (bias + pos-emb are summed before instead of after the accumulator)
    from mirage_b200 import functional as Fn
    specs = [model.input_adapters[d].visible_spec(x[d]) for d in MODS]
    assert all(s is not None for s in specs)
    with torch.no_grad():
        tok = Fn.embed_visible([s[0] for s in specs], [t for s in specs for t in s[1]], masks[1], model.global_tokens)
        all_tok, _ = model._embed_all(x)
        ref = Fn.token_gather(all_tok, masks[1], model.global_tokens[0]).reshape(tok.shape)
    assert (tok - ref).abs().max().item() <= 2e-6 * ref.abs().max().item()

    out = {}
    for vis in (True, False):
        model.visible_embedding = vis
        preds, losses, grads = b200_step(model, crits, x, masks)
        out[vis] = ({k: v.detach().float().clone() for k, v in preds.items()}, losses,
                    {k: v.detach().float().clone() for k, v in grads.items()})
    for d in MODS:
        a, b = out[True][0][d], out[False][0][d]
        # identical tokens up to 1 ulp, then 2-12 bf16 blocks: the repo's standard tolerances (tests/helpers.py)
        assert (a - b).abs().max().item() <= 2e-2 * b.abs().max().item(), d
        assert abs(out[True][1][d] - out[False][1][d]) <= 2e-3 * max(1.0, abs(out[False][1][d]))
    assert set(out[True][2]) == set(out[False][2])
    for k, gb in out[False][2].items():
        ga = out[True][2][k]
        den = gb.norm().item()
        if den == 0.0:
            assert ga.norm().item() == 0.0, k
            continue
        assert (ga - gb).norm().item() <= 5e-2 * den, (k, (ga - gb).norm().item() / den)


def test_ragged_modalities_no_grad():
    """Two image modalities with different token counts (512 x 512 -> 256 tokens, 512 x 256 -> 128, pos-emb resized),
    two global tokens, inference (no bf16 twins are built): kept-token embedding == embed-all-then-gather."""
    from mirage_b200 import functional as Fn
    from mirage_b200.input_adapters import PatchedInputAdapter
    from mirage_b200.model import MIRAGEModel
    from pretrain_case import pretrain_args
    dev = torch.device("cuda:0")
    mods = ["bscan", "slo"]
    ins = {d: PatchedInputAdapter(num_channels=1, stride_level=1, patch_size_full=(32, 32), image_size=(512, 512))
           for d in mods}
    model = MIRAGEModel(pretrain_args(mods), input_adapters=ins, output_adapters=None, num_global_tokens=2,
                        dim_tokens=128, depth=1, num_heads=2, drop_path_rate=0.0)
    load_synth(model, seed=9)
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(4)
    B, n_keep = 5, 61
    x = {"bscan": torch.rand(B, 1, 512, 512, generator=g).to(dev), "slo": torch.rand(B, 1, 512, 256, generator=g).to(dev)}
    n_all = 256 + 128
    ids_keep = torch.stack([torch.randperm(n_all, generator=g)[:n_keep] for _ in range(B)]).to(dev)
    specs = [model.input_adapters[d].visible_spec(x[d]) for d in mods]
    assert all(s is not None for s in specs) and [s[0]['count'] for s in specs] == [256, 128]
    with torch.no_grad():
        tok = Fn.embed_visible([s[0] for s in specs], [t for s in specs for t in s[1]], ids_keep, model.global_tokens)
        all_tok, counts = model._embed_all(x)
        ref = Fn.token_gather(all_tok, ids_keep, model.global_tokens[0]).reshape(tok.shape)
    assert list(counts.values()) == [256, 128]
    assert (tok - ref).abs().max().item() <= 2e-6 * ref.abs().max().item()
    # and through MIRAGEModel.forward with the masks injected (both orders, encoder tokens)
    mask_all = torch.ones(B, n_all, dtype=torch.long, device=dev)
    mask_all.scatter_(1, ids_keep, 0)
    ids_restore = torch.argsort(torch.cat([ids_keep, torch.zeros(B, 0, dtype=torch.long, device=dev)], 1), dim=1)
    tm = dict(zip(mods, torch.split(mask_all, [256, 128], dim=1)))
    model.generate_random_masks = lambda *a, **k: (tm, ids_keep, ids_restore)
    outs = []
    for vis in (True, False):
        model.visible_embedding = vis
        with torch.no_grad():
            enc, _ = model(x, num_encoded_tokens=n_keep)
        outs.append(enc.float())
    assert outs[0].shape == (B, n_keep + 2, 128)
    assert (outs[0] - outs[1]).abs().max().item() <= 2e-2 * outs[1].abs().max().item()
