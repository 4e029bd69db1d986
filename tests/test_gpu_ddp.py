"""-m gpu: the data-parallel pretraining step with the REAL model on CUDA, two processes sharing
cuda:0 over gloo (NCCL refuses two ranks on one device; the driver's GPU tier has one GPU).  Exercises
the gradient sink (kernels accumulating straight into the buckets), the hook/done accounting and the
bucket launch order: averaged gradients must equal the single-process gradients on the concatenated
batch with the same masks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup(batch, dev=None):
    from helpers import load_synth, synth_images
    from pretrain_case import MODS, build_criteria, build_pretrain_model, sample_masks
    dev = dev or torch.device("cuda:0")
    model, _ = build_pretrain_model("tiny")
    load_synth(model, seed=3)
    model = model.to(dev).train()
    x = {k: v.to(dev) for k, v in synth_images(batch, MODS, seed=31).items()}
    tm, keep, restore = sample_masks(model, batch, 98, seed=17)
    masks = ({k: v.to(dev) for k, v in tm.items()}, keep.to(dev), restore.to(dev))
    return model, build_criteria(), x, masks, MODS


def _run(model, crits, x, masks, mods, sl):
    tm, keep, restore = masks
    model.generate_random_masks = lambda *a, **k: ({d: v[sl] for d, v in tm.items()}, keep[sl], restore[sl])
    xs = {k: v[sl] for k, v in x.items()}
    preds, m = model(xs, num_encoded_tokens=98, alphas=1.0, sample_tasks_uniformly=False)
    sum(crits[d](preds[d].float(), xs[d], mask=m[d]) for d in mods).backward()


def _worker(rank, world, port, ret, backend="gloo"):
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    for q in (str(root), str(root / "tests")):
        if q not in sys.path:
            sys.path.insert(0, q)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from mirage_b200 import functional as Fn
    from mirage_b200.ddp import GradBucketAllReduce
    model, crits, x, masks, mods = _setup(4, dev)
    ddp = GradBucketAllReduce(model, bucket_mb=0.5)
    assert Fn._grad_sink is ddp and len(ddp.buckets) > 3
    for _ in range(2):
        ddp.zero_grad()
        _run(model, crits, x, masks, mods, slice(rank * 2, rank * 2 + 2))
        ddp.finish()
    torch.cuda.synchronize()
    assert all(b.pending >= 0 for b in ddp.buckets)
    if rank == 0:
        ret["grads"] = {n: p.grad.detach().float().cpu() for n, p in model.named_parameters() if p.requires_grad}
    dist.barrier()
    dist.destroy_process_group()


def _check_two_rank(backend):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret, backend), nprocs=2, join=True)
    got = ret["grads"]
    # single process, no sink: per-rank losses are means over 2 samples each, so the rank average equals
    # the mean of the two half-batch gradients
    model, crits, x, masks, mods = _setup(4)
    ref = None
    for r in range(2):
        model.zero_grad(set_to_none=True)
        _run(model, crits, x, masks, mods, slice(r * 2, r * 2 + 2))
        g = {n: p.grad.detach().float().cpu() for n, p in model.named_parameters() if p.grad is not None}
        ref = g if ref is None else {n: ref[n] + g[n] for n in g}
    ref = {n: v / 2 for n, v in ref.items()}
    big = max(v.abs().max().item() for v in ref.values())
    for n, v in ref.items():
        tol = 2e-3 * v.abs().max().item() + 1e-6 * big
        assert (got[n] - v).abs().max().item() <= tol, (n, (got[n] - v).abs().max().item(), tol)


def test_two_rank_step_matches_single_process():
    _check_two_rank("gloo")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="the NCCL gradient path needs two GPUs")
def test_two_gpu_nccl_step_matches_single_process():
    """The real exchange: one process per GPU, ncclAvg all-reduce of the flat fp32 buckets launched from
    inside backward (ddp.py _launch); averaged gradients == single-process gradients."""
    _check_two_rank("nccl")
