"""world_size-2 gloo test of the gradient-exchange host logic (runs on CPU): bucketed, hook-driven
mean all-reduce must equal single-process gradients on the concatenated batch."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(16, 64), nn.GELU(), nn.LayerNorm(64), nn.Linear(64, 64), nn.GELU(),
                         nn.Linear(64, 4))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mirage_b200.ddp import GradBucketAllReduce
    model = _model()
    model[0].bias.requires_grad_(False)                     # a frozen parameter must be skipped
    ddp = GradBucketAllReduce(model, bucket_mb=0.01)        # tiny buckets -> several of them
    assert len(ddp.buckets) > 2
    g = torch.Generator().manual_seed(42)
    x = torch.randn(8, 16, generator=g)
    y = torch.randn(8, 4, generator=g)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    for _ in range(2):                                       # second iteration exercises zero_grad()
        ddp.zero_grad()
        loss = ((model(xs) - ys) ** 2).mean()
        loss.backward()
        ddp.finish()
    grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.requires_grad}
    norm = ddp.grad_norm().item()
    if rank == 0:
        ret["grads"] = grads
        ret["norm"] = norm
    gathered = [None, None]
    dist.all_gather_object(gathered, norm)
    assert abs(gathered[0] - gathered[1]) < 1e-7            # identical on every rank
    dist.destroy_process_group()


def test_bucketed_allreduce_matches_single_process():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    model = _model()
    model[0].bias.requires_grad_(False)
    g = torch.Generator().manual_seed(42)
    x = torch.randn(8, 16, generator=g)
    y = torch.randn(8, 4, generator=g)
    # mean over ranks of per-rank mean losses == mean loss over the concatenated batch
    ((model(x) - y) ** 2).mean().backward()
    ref = {n: p.grad for n, p in model.named_parameters() if p.requires_grad}
    got = ret["grads"]
    assert set(ref) == set(got) and "0.bias" not in got
    for n in ref:
        assert torch.allclose(got[n], ref[n], rtol=1e-5, atol=1e-7), n
    total = torch.sqrt(sum((v ** 2).sum() for v in ref.values())).item()
    assert abs(ret["norm"] - total) <= 1e-5 * total


def test_bucket_views_are_128_byte_aligned():
    """Every gradient view starts on a 128-byte boundary of its bucket (the CUDA kernels write them with
    16-byte vector stores / vector reductions) -- also behind a 5-element bias -- and zero_grad() keeps
    the views in place."""
    from mirage_b200.ddp import GradBucketAllReduce
    torch.manual_seed(1)
    model = nn.Sequential(nn.Linear(7, 5), nn.Linear(5, 33), nn.LayerNorm(33), nn.Linear(33, 3))
    ddp = GradBucketAllReduce(model, bucket_mb=0.001)
    assert len(ddp.buckets) >= 2
    seen = set()
    for b in ddp.buckets:
        base = b.flat.data_ptr()
        for p in b.params:
            off = p.grad.data_ptr() - base
            assert off % 128 == 0 and 0 <= off < b.flat.numel() * 4, (tuple(p.shape), off)
            assert p.grad.shape == p.shape and p.grad.is_contiguous()
            seen.add(id(p))
    assert seen == {id(p) for p in model.parameters()}
    ptrs = [p.grad.data_ptr() for p in model.parameters()]
    model(torch.randn(4, 7)).sum().backward()
    ddp.finish()
    g0 = [p.grad.clone() for p in model.parameters()]
    assert all(float(g.abs().sum()) > 0 for g in g0[:2])
    ddp.zero_grad()
    assert ptrs == [p.grad.data_ptr() for p in model.parameters()]
    assert all(float(p.grad.abs().sum()) == 0.0 for p in model.parameters())
    assert ddp.num_params == sum(p.numel() for p in model.parameters())
