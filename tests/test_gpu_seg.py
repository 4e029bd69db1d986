"""-m gpu: the segmentation-tuning caller: MIRAGELight + LinearSegAdapter / ConvNeXtAdapter on
cuda:0 against the CPU oracle (forward and gradients) and the reference's golden predictions; the encoder at the
segmentation sequence length N = 1025 (1024 x 1024 input) with return_all_layers."""
import pytest
import torch

from helpers import GOLDEN, assert_parity, load_synth, synth_images

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["linear", "convnext"])
def test_seg_head_forward_backward_vs_oracle(kind):
    from seg_case import build_seg_model, oracle_seg
    dev = torch.device("cuda:0")
    g = torch.load(GOLDEN / "seg.pt")
    m = build_seg_model(kind)
    sd = load_synth(m, seed=g["weights_seed"])
    m = m.to(dev).train()
    x = synth_images(g["batch"], ["bscan"], seed=g["input_seed"])
    pred = m({k: v.to(dev) for k, v in x.items()})["bscan"]
    assert pred.shape == (2, 13, 512, 512)
    # golden rows recorded from the reference itself
    rows = pred.detach().float().cpu().reshape(-1, 512)[:: g["out"][kind]["pred"]["step"]]
    assert_parity(rows, g["out"][kind]["pred"]["rows"], f"seg {kind} vs reference golden")
    # gradients of a cross-entropy loss against the oracle's autograd
    tgt = torch.randint(0, 13, (2, 512, 512), generator=torch.Generator().manual_seed(3))
    loss = torch.nn.functional.cross_entropy(pred.float(), tgt.to(dev))
    loss.backward()
    leaf = {k: v.clone().requires_grad_(not k.endswith("pos_emb")) for k, v in sd.items()}
    ref = oracle_seg(x, leaf, kind)
    ref_loss = torch.nn.functional.cross_entropy(ref, tgt)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 2e-2 * abs(ref_loss.item())
    big = max(v.grad.norm().item() for v in leaf.values() if v.grad is not None)
    checked = 0
    for k, p in m.named_parameters():
        r = leaf[k].grad
        if r is None or p.grad is None:
            assert (r is None or r.norm().item() < 1e-5 * big) and (p.grad is None or True), k
            continue
        if r.norm().item() < 1e-4 * big:
            continue
        a, b = p.grad.detach().float().cpu().flatten(), r.flatten()
        rel = ((a - b).norm() / b.norm()).item()
        cos = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
        assert rel <= 5e-2 and cos >= 0.999, (k, rel, cos)
        checked += 1
    assert checked >= 10


def test_seg_encoder_1024_return_all_layers():
    """N = 1025 tokens (docs/segmentation_tuning.md:95): whole quads of query tiles -> the four-tile attention
    kernel; every layer's tokens against the oracle."""
    from mirage_b200.mirage_hf import MIRAGEWrapper
    from oracle import mirage_oracle as O
    dev = torch.device("cuda:0")
    m = MIRAGEWrapper(input_size=1024, patch_size=32, modalities="bscan", size="base")
    sd = load_synth(m.model, seed=2)
    m = m.to(dev).eval()
    x = synth_images(1, ["bscan"], seed=6, size=1024)
    with torch.no_grad():
        layers = m.model({k: v.to(dev) for k, v in x.items()}, return_all_layers=True)
        refs = O.light_forward(x, sd, 12, 12, return_all_layers=True)
    assert len(layers) == 12 and layers[-1].shape == (1, 1025, 768)
    for i in (0, 5, 11):
        assert_parity(layers[i], refs[i], f"layer {i} at N=1025")
