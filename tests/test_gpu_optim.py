"""-m gpu: optim.FusedAdamW (csrc/optim.cu) against torch.optim.AdamW in fp32 -- the optimizer the reference
builds in mutils/optim_factory.py:171-172 -- incl. parameter groups, clipping / skipping semantics of
mutils/native_scaler.py:16-37, the bf16 weight twins and checkpoint compatibility."""
import copy

import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu

REL = 1e-6   # stated tolerance: <= 1e-6 relative to the parameter scale over 3 steps


def _toy(dev, seed=0):
    torch.manual_seed(seed)
    m = nn.Sequential(nn.Linear(96, 200), nn.LayerNorm(200), nn.Linear(200, 4099, bias=True), nn.Linear(4099, 7))
    return m.to(dev)


def _groups(model, wd=0.05):
    decay = [p for n, p in model.named_parameters() if p.ndim > 1]
    no_decay = [p for n, p in model.named_parameters() if p.ndim <= 1]
    return [{"params": decay, "weight_decay": wd, "lr_scale": 1.0}, {"params": no_decay, "weight_decay": 0.0, "lr_scale": 0.5}]


def _set_grads(model, step, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(100 + step)
    for p in model.parameters():
        p.grad = torch.randn(p.shape, device=p.device, generator=g) * scale


def _max_rel(a_model, b_model):
    worst = 0.0
    for (n, a), (_, b) in zip(a_model.named_parameters(), b_model.named_parameters()):
        worst = max(worst, ((a - b).abs().max() / (b.abs().max() + 1e-12)).item())
    return worst


@pytest.mark.parametrize("mode", ["plain", "clip", "skip"])
def test_fused_adamw_matches_torch(mode):
    from mirage_b200.optim import FusedAdamW
    dev = torch.device("cuda:0")
    ref_m, our_m = _toy(dev), _toy(dev)
    ref = torch.optim.AdamW(_groups(ref_m), lr=1e-3, betas=(0.9, 0.95), eps=1e-8)
    our = FusedAdamW(_groups(our_m), lr=1e-3, betas=(0.9, 0.95), eps=1e-8)
    lrs = [1e-3, 7e-4, 2e-3, 1e-3]
    for step in range(4):
        for opt in (ref, our):
            for g in opt.param_groups:      # the per-step assignment of run_pretraining.py:683-688
                g["lr"] = lrs[step] * g["lr_scale"]
        scale = 30.0 if (mode == "skip" and step == 1) else 1.0
        _set_grads(ref_m, step, scale)
        _set_grads(our_m, step, scale)
        params = list(ref_m.parameters())
        if mode == "clip":
            ref_norm = torch.nn.utils.clip_grad_norm_(params, 5.0)
            ref.step()
            norm = our.step(clip_grad=5.0)
        elif mode == "skip":
            ref_norm = torch.linalg.vector_norm(torch.stack([p.grad.norm() for p in params]))
            if ref_norm < 2000.0:
                ref.step()
            norm = our.step(skip_grad=2000.0)
            assert int(our.last_step_skipped.item()) == (0 if ref_norm < 2000.0 else 1)
        else:
            ref_norm = torch.linalg.vector_norm(torch.stack([p.grad.norm() for p in params]))
            ref.step()
            norm = our.step()
        assert abs(norm.item() - ref_norm.item()) <= 1e-5 * ref_norm.item()
        assert _max_rel(our_m, ref_m) <= REL, (mode, step, _max_rel(our_m, ref_m))
    expected_steps = 3 if mode == "skip" else 4
    assert int(our.step_count.item()) == expected_steps
    for (pa, pb) in zip(our_m.parameters(), ref_m.parameters()):
        sa, sb = our.state[pa], ref.state[pb]
        for key in ("exp_avg", "exp_avg_sq"):   # same tolerance as the parameters: relative to the tensor's scale
            err = (sa[key] - sb[key]).abs().max().item()
            assert err <= 2e-6 * sb[key].abs().max().item(), (key, err)


def test_fused_adamw_refreshes_bf16_twins_and_zeroes_grads():
    """The GEMM operand twins follow the parameters without any cast kernel, and zero_grad_in_step leaves the
    gradient buffers zeroed for the next backward."""
    from mirage_b200 import functional as Fn
    from mirage_b200 import ops
    from mirage_b200.optim import FusedAdamW
    dev = torch.device("cuda:0")
    m = _toy(dev, seed=3)
    w = m[2].weight
    twin = Fn.bf16_weight(w)
    ptr = twin.data_ptr()
    opt = FusedAdamW(_groups(m), lr=1e-2, betas=(0.9, 0.95), zero_grad_in_step=True)
    _set_grads(m, 0)
    grads = [p.grad for p in m.parameters()]
    ops.reset_launch_count()
    opt.step()
    assert ops.launch_count() == 3                      # prepare + update + finish; no cast kernels
    again = Fn.bf16_weight(w)
    assert again.data_ptr() == ptr and ops.launch_count() == 3
    assert torch.equal(again, w.detach().to(torch.bfloat16))
    assert all(float(g.abs().max()) == 0.0 for g in grads)
    # an eager modification through PyTorch is noticed (version counter) and re-cast into the SAME buffer
    with torch.no_grad():
        w.mul_(0.5)
    again = Fn.bf16_weight(w)
    assert again.data_ptr() == ptr and torch.equal(again, w.detach().to(torch.bfloat16))


def test_fused_adamw_state_dict_round_trips_with_torch():
    """Checkpoints written by the reference (torch.optim.AdamW state, mutils/checkpoint.py:9-25) load into
    FusedAdamW and vice versa; training continues identically."""
    from mirage_b200.optim import FusedAdamW
    dev = torch.device("cuda:0")
    ref_m, our_m = _toy(dev, 5), _toy(dev, 5)
    ref = torch.optim.AdamW(_groups(ref_m), lr=1e-3, betas=(0.9, 0.95))
    for step in range(2):
        _set_grads(ref_m, step)
        ref.step()
    our_m.load_state_dict(ref_m.state_dict())
    our = FusedAdamW(_groups(our_m), lr=1e-3, betas=(0.9, 0.95))
    our.load_state_dict(copy.deepcopy(ref.state_dict()))
    _set_grads(ref_m, 2)
    _set_grads(our_m, 2)
    ref.step()
    our.step()
    assert int(our.step_count.item()) == 3
    assert _max_rel(our_m, ref_m) <= REL
    sd = our.state_dict()
    ref2 = torch.optim.AdamW(_groups(ref_m), lr=1e-3, betas=(0.9, 0.95))
    ref2.load_state_dict(sd)
    assert all(float(s["step"]) == 3.0 for s in ref2.state.values())


def test_fused_adamw_inside_cuda_graph_follows_host_schedule():
    from mirage_b200.optim import FusedAdamW
    dev = torch.device("cuda:0")
    ref_m, our_m = _toy(dev, 7), _toy(dev, 7)
    ref = torch.optim.AdamW(_groups(ref_m), lr=1e-3, betas=(0.9, 0.95))
    our = FusedAdamW(_groups(our_m), lr=1e-3, betas=(0.9, 0.95))
    fixed = [torch.zeros_like(p) for p in our_m.parameters()]
    for p, g in zip(our_m.parameters(), fixed):
        p.grad = g
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        our.step()                      # builds the tables (zero gradients: parameters only decay)
    torch.cuda.current_stream().wait_stream(side)
    ref.zero_grad(set_to_none=False)
    for p in ref_m.parameters():
        p.grad = torch.zeros_like(p)
    ref.step()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        our.step()
    # capture does not execute: both sides are one step in
    for step, lr in enumerate([5e-4, 2e-3, 1e-3]):
        _set_grads(ref_m, step)
        g = torch.Generator(device="cuda").manual_seed(100 + step)
        for buf in fixed:
            buf.copy_(torch.randn(buf.shape, device=dev, generator=g))
        for opt in (ref, our):
            for grp in opt.param_groups:
                grp["lr"] = lr * grp["lr_scale"]
        ref.step()
        our.refresh_hyper()
        graph.replay()
        assert _max_rel(our_m, ref_m) <= REL, (step, _max_rel(our_m, ref_m))
    assert int(our.step_count.item()) == 4
