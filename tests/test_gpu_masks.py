"""-m gpu: token-mask sampling.  (1) the reference op sequence on the CUDA generator keeps the reference's
invariants; (2) the fused on-device sampler (csrc/masks.cu) keeps them too and draws from
the same distribution as mirage/model.py:168-239 (oracle: oracle.mirage_oracle.random_masks, pinned
bit-exactly to the reference by tests/golden/masks.pt)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

MODS = ["bscan", "slo", "bscanlayermap"]


def _model(dev):
    from pretrain_case import build_pretrain_model
    model, _ = build_pretrain_model("tiny")
    return model.to(dev)


def _check_invariants(tm, keep, restore, B, n_vis, counts):
    n_all = sum(counts)
    mask_all = torch.cat([tm[d] for d in tm], dim=1)
    assert mask_all.dtype == keep.dtype == restore.dtype == torch.int64
    assert mask_all.shape == (B, n_all) and keep.shape == (B, n_vis) and restore.shape == (B, n_all)
    assert set(mask_all.unique().tolist()) <= {0, 1}
    # exactly n_vis visible tokens per sample
    assert torch.equal((mask_all == 0).sum(1), torch.full((B,), n_vis, device=keep.device))
    # ids_restore is a permutation and inverts the shuffle ids_keep is a prefix of
    ar = torch.arange(n_all, device=keep.device).expand(B, -1)
    assert torch.equal(restore.sort(dim=1).values, ar)
    shuffle = torch.argsort(restore, dim=1)
    assert torch.equal(shuffle[:, :n_vis], keep)
    assert torch.equal(torch.gather(restore, 1, keep), ar[:, :n_vis])
    # masks consistent with the ids: visible <=> position < n_vis
    assert torch.equal(mask_all, (restore >= n_vis).long())
    assert torch.equal(torch.gather(mask_all, 1, keep), torch.zeros_like(keep))


@pytest.mark.parametrize("B,n_vis", [(16, 98), (5, 1), (3, 768)])
def test_reference_sampler_on_cuda_generator_invariants(B, n_vis):
    dev = torch.device("cuda:0")
    model = _model(dev)
    torch.manual_seed(5)
    toks = {d: torch.empty(B, 256, 0, device=dev) for d in MODS}
    tm, keep, restore = model.generate_random_masks(toks, n_vis, alphas=1.0)
    _check_invariants(tm, keep, restore, B, n_vis, [256, 256, 256])


@pytest.mark.parametrize("B,n_vis,counts,uniform", [(64, 98, [256, 256, 256], False), (7, 1, [256, 256, 256], False),
                                                    (4, 768, [256, 256, 256], False), (9, 100, [256, 64], True),
                                                    (3, 200, [1024, 1024, 1024], False), (32, 98, [256, 256, 256], True)])
def test_device_sampler_invariants(B, n_vis, counts, uniform):
    dev = torch.device("cuda:0")
    model = _model(dev).set_mask_sampler("device", seed=1234)
    toks = {f"t{i}": torch.empty(B, n, 0, device=dev) for i, n in enumerate(counts)}
    tm, keep, restore = model.generate_random_masks(toks, n_vis, alphas=1.0, sample_tasks_uniformly=uniform)
    _check_invariants(tm, keep, restore, B, n_vis, counts)
    # same seed + same draw index -> same masks; next draw differs (the kernel advances its counter)
    model2 = _model(dev).set_mask_sampler("device", seed=1234)
    tm2, keep2, restore2 = model2.generate_random_masks(toks, n_vis, alphas=1.0, sample_tasks_uniformly=uniform)
    assert torch.equal(restore, restore2) and torch.equal(keep, keep2)
    assert int(model._mask_rng["draw"].item()) == 1
    tm3, keep3, restore3 = model.generate_random_masks(toks, n_vis, alphas=1.0, sample_tasks_uniformly=uniform)
    _check_invariants(tm3, keep3, restore3, B, n_vis, counts)
    if n_vis < sum(counts) or B > 1:
        assert not torch.equal(restore, restore3)


def test_device_sampler_matches_reference_distribution():
    """Per-task visible counts and token-level visibility frequencies: two-sample Kolmogorov-Smirnov tests
    against the reference sampler (alpha = 1 and alpha = 0.3), 4096 samples each."""
    from scipy import stats

    from oracle import mirage_oracle as O
    dev = torch.device("cuda:0")
    counts, n_vis, B = [256, 256, 256], 98, 4096
    for alpha in (1.0, 0.3):
        model = _model(dev).set_mask_sampler("device", seed=99)
        toks = {d: torch.empty(B, 256, 0, device=dev) for d in MODS}
        tm, keep, restore = model.generate_random_masks(toks, n_vis, alphas=alpha)
        ours = torch.stack([(tm[d] == 0).sum(1) for d in MODS], 1).cpu().numpy()
        torch.manual_seed(17)
        rtm, rkeep, rrestore = O.random_masks(counts, B, n_vis, alphas=alpha)
        ref = torch.stack([(m == 0).sum(1) for m in rtm], 1).numpy()
        for t in range(3):
            ks = stats.ks_2samp(ours[:, t], ref[:, t])
            assert ks.pvalue > 1e-3, (alpha, t, ks)
        assert abs(ours.mean() - n_vis / 3) < 1e-6
        # every token of a task is visible equally often: mean 98/768, and the position of a token in ids_keep
        # is uniform (chi-square on coarse bins)
        vis = torch.cat([tm[d] for d in MODS], 1).eq(0).float().mean(0).cpu()
        assert abs(vis.mean().item() - n_vis / 768) < 1e-6
        assert vis.std().item() < 3.0 * (n_vis / 768 * (1 - n_vis / 768) / B) ** 0.5 * 1.5
        pos = keep.cpu().numpy()
        hist = [((pos[:, j] // 96)[:, None] == range(8)).sum(0) for j in (0, 50, 97)]
        for h in hist:
            assert stats.chisquare(h).pvalue > 1e-4, h


def test_device_sampler_in_cuda_graph_draws_fresh_masks():
    dev = torch.device("cuda:0")
    model = _model(dev).set_mask_sampler("device", seed=7)
    toks = {d: torch.empty(8, 256, 0, device=dev) for d in MODS}
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        model.generate_random_masks(toks, 98, alphas=1.0)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        tm, keep, restore = model.generate_random_masks(toks, 98, alphas=1.0)
    seen = []
    for _ in range(3):
        g.replay()
        _check_invariants(tm, keep, restore, 8, 98, [256, 256, 256])
        seen.append(restore.clone())
    assert not torch.equal(seen[0], seen[1]) and not torch.equal(seen[1], seen[2])
    assert int(model._mask_rng["draw"].item()) == 4


def test_pretrain_step_runs_with_device_sampler():
    from helpers import load_synth, synth_images
    from pretrain_case import build_criteria, build_pretrain_model
    dev = torch.device("cuda:0")
    model, _ = build_pretrain_model("tiny")
    load_synth(model, seed=3)
    model = model.to(dev).train().set_mask_sampler("device", seed=3)
    crits = build_criteria()
    x = {k: v.to(dev) for k, v in synth_images(4, MODS, seed=2).items()}
    preds, masks = model(x, num_encoded_tokens=98, alphas=1.0, sample_tasks_uniformly=False)
    loss = sum(crits[d](preds[d].float(), x[d], mask=masks[d]) for d in MODS)
    loss.backward()
    assert torch.isfinite(loss) and loss.item() > 0
    assert sum(int((masks[d] == 0).sum()) for d in MODS) == 4 * 98
