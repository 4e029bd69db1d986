"""-m gpu: fused LayerNorm + token mean-pool (csrc/pool.cu) against the reference's op
sequence norm -> pool (mirage_wrapper.py:217-244) in fp32 torch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, g, b, eps, ranges):
    y = torch.nn.functional.layer_norm(x, (x.shape[-1],), g, b, eps)
    return torch.cat([y[:, r0:r1].mean(1) for r0, r1 in ranges], 1)


@pytest.mark.parametrize("B,N,D,ranges", [(64, 257, 1024, [(0, 256)]), (3, 257, 768, [(256, 257)]),
                                           (5, 513, 1024, [(0, 512), (512, 513)]), (2, 99, 128, [(0, 98)]),
                                           (1, 10, 256, [(0, 9), (9, 10)])])
def test_ln_meanpool_forward_backward(B, N, D, ranges):
    from mirage_b200 import functional as Fn
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + N)
    x = (torch.randn(B, N, D, device=dev, generator=g) * 2 + 0.5).requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(D, device=dev, generator=g)).requires_grad_(True)
    beta = (0.1 * torch.randn(D, device=dev, generator=g)).requires_grad_(True)
    w = torch.randn(B, len(ranges) * D, device=dev, generator=g)
    out = Fn.ln_meanpool(x, gamma, beta, 1e-6, ranges)
    ref = _ref(x, gamma, beta, 1e-6, ranges)
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-6
    got = torch.autograd.grad((out * w).sum(), [x, gamma, beta])
    exp = torch.autograd.grad((ref * w).sum(), [x, gamma, beta])
    for a, e, name in zip(got, exp, ("dx", "dgamma", "dbeta")):
        assert (a - e).abs().max().item() <= 1e-4 * e.abs().max().item() + 1e-7, name
    # rows outside every range get an exact zero gradient
    covered = torch.zeros(N, dtype=torch.bool)
    for r0, r1 in ranges:
        covered[r0:r1] = True
    if (~covered).any():
        assert float(got[0][:, ~covered.to(dev)].abs().max()) == 0.0


def test_cls_wrapper_uses_fused_pool():
    """The three pooling variants route through ln_meanpool (and a user-overridden pool() does not)."""
    from cls_case import build_cls_model
    from mirage_b200 import ops
    dev = torch.device("cuda:0")
    for pool in ("global", "cls", "token_mix"):
        m, _ = build_cls_model(pool, 21, device=dev)
        assert m._pool_ranges(257) is not None
        x = torch.rand(2, 1, 512, 512, device=dev)
        seen = []
        ops.set_recorder(lambda name, work, unit: (seen.append(name), ops._NULL)[1])
        try:
            with torch.no_grad():
                fused = m(x)
        finally:
            ops.set_recorder(None)
        assert "ln_meanpool_fwd" in seen
        # reference op order on the same encoder output: LayerNorm -> pool -> head
        with torch.no_grad():
            torch.manual_seed(0)
            tok, _ = m.model({"bscan": x}, mask_inputs=False)
            y = torch.nn.functional.layer_norm(tok, (tok.shape[-1],), m.norm.weight, m.norm.bias, m.norm.eps)
            ref = m.head(m.pool(y))
        assert (fused - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()

    class Custom(type(m)):
        def pool(self, x):
            return x[:, :-1].amax(dim=1)
    m.__class__ = Custom          # a user subclass with its own pooling falls back to LayerNorm + pool()
    assert m._pool_ranges(257) is None
