"""Pretraining workloads for bench.py (cfg 3 / cfg 4 of BASELINE.json): model + criteria + AdamW +
gradient all-reduce on the GPU arm, and the oracle's forward+backward on the CPU arm."""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "tests"))


def build_pretrain_step(size, mods, per_gpu, dev, rank, world):
    from helpers import load_synth, synth_images
    from mirage_b200.ddp import GradBucketAllReduce
    from pretrain_case import build_criteria, build_pretrain_model

    model, _ = build_pretrain_model(size, mods)
    load_synth(model, seed=3)
    model = model.to(dev).train()
    crits = build_criteria(mods)
    ddp = GradBucketAllReduce(model, bucket_mb=64)
    decay, no_decay = [], []
    skip = model.no_weight_decay()
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (no_decay if (p.ndim <= 1 or n.endswith(".bias") or n in skip) else decay).append(p)
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": 0.05}, {"params": no_decay, "weight_decay": 0.0}],
                            lr=1e-4 * per_gpu * world / 256, betas=(0.9, 0.95), fused=True, capturable=True)

    base = synth_images(8, mods, seed=1234 + rank)
    reps = per_gpu // 8 + 1
    host_in = {k: v.repeat(reps, *([1] * (v.dim() - 1)))[:per_gpu].contiguous().pin_memory() for k, v in base.items()}
    dev_in = {k: v.to(dev) for k, v in host_in.items()}
    torch.manual_seed(100 + rank)
    host_loss = torch.zeros(1).pin_memory()

    def one_step(x):
        ddp.zero_grad()
        preds, masks = model(x, num_encoded_tokens=98, alphas=1.0, sample_tasks_uniformly=False)
        loss = sum(crits[d](preds[d].float(), x[d], mask=masks[d]) for d in mods)
        loss.backward()
        ddp.finish()
        opt.step()
        return loss

    def step():
        return one_step(dev_in)

    def step_e2e():
        x = {k: v.to(dev, non_blocking=True) for k, v in host_in.items()}
        loss = one_step(x)
        host_loss.copy_(loss.detach().reshape(1), non_blocking=True)
        return loss

    def graph_hooks():
        """Whole-step CUDA graph (single rank): forward + losses + backward + AdamW are captured once;
        the token masks are sampled eagerly every step exactly as the reference does (CPU Dirichlet +
        device noise, mirage/model.py:168-239) and copied into the fixed tensors the graph reads."""
        from mirage_b200.graphs import GraphedCallable
        sampler = model.generate_random_masks
        toks = {d: torch.empty(per_gpu, 256, 0, device=dev) for d in mods}
        tm0, keep0, restore0 = sampler(toks, 98, alphas=1.0)
        fixed = ({d: t.clone() for d, t in tm0.items()}, keep0.clone(), restore0.clone())
        x_fixed = {k: v.clone() for k, v in dev_in.items()}

        def resample():
            tm, keep, restore = sampler(toks, 98, alphas=1.0)
            for d in mods:
                fixed[0][d].copy_(tm[d])
            fixed[1].copy_(keep)
            fixed[2].copy_(restore)

        model.generate_random_masks = lambda *a, **k: fixed
        g = GraphedCallable(lambda: one_step(x_fixed)).capture()

        def gstep():
            resample()
            return g()

        # end-to-end leg: the batch of step k+1 travels host -> device (pinned memory, side stream) while
        # step k computes, as a prefetching data loader does; each step begins with a device-to-device copy
        # of the staged batch into the tensors the graph reads
        side = torch.cuda.Stream(dev)
        staging = {k: torch.empty_like(v) for k, v in dev_in.items()}
        state = {"ready": None}

        def prefetch(after):
            with torch.cuda.stream(side):
                if after is not None:
                    side.wait_event(after)
                for k in staging:
                    staging[k].copy_(host_in[k], non_blocking=True)
                state["ready"] = side.record_event()

        def gstep_e2e():
            main = torch.cuda.current_stream(dev)
            if state["ready"] is None:
                prefetch(None)
            main.wait_event(state["ready"])
            for k in x_fixed:
                x_fixed[k].copy_(staging[k], non_blocking=True)
            prefetch(main.record_event())          # next step's batch, overlapped with this step
            resample()
            loss = g()
            host_loss.copy_(loss.detach().reshape(1), non_blocking=True)
            return loss
        return gstep, gstep_e2e

    h2d = sum(v.numel() * v.element_size() for v in host_in.values())
    return step, step_e2e, h2d, 4, graph_hooks


def build_cls_step(size, per_gpu, dev, rank, world):
    """cfg 5: miragecls_factory['global'] full fine-tune step (forward + CE + backward + gradient
    exchange + AdamW), bscan 512x512, 5 classes."""
    from cls_case import build_cls_model
    from helpers import synth_images
    from mirage_b200.ddp import GradBucketAllReduce
    model, _ = build_cls_model("global", 21, device=dev, size=size)
    model.train()
    ddp = GradBucketAllReduce(model, bucket_mb=64)
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.05,
                            fused=True)
    base = synth_images(8, ["bscan"], seed=1234 + rank)["bscan"]
    host_x = base.repeat(per_gpu // 8 + 1, 1, 1, 1)[:per_gpu].contiguous().pin_memory()
    dev_x = host_x.to(dev)
    tgt = torch.randint(0, 5, (per_gpu,), generator=torch.Generator().manual_seed(5 + rank)).to(dev)
    torch.manual_seed(100 + rank)
    host_loss = torch.zeros(1).pin_memory()

    def one_step(x):
        ddp.zero_grad()
        loss = torch.nn.functional.cross_entropy(model(x).float(), tgt)
        loss.backward()
        ddp.finish()
        opt.step()
        return loss

    def step():
        return one_step(dev_x)

    def step_e2e():
        loss = one_step(host_x.to(dev, non_blocking=True))
        host_loss.copy_(loss.detach().reshape(1), non_blocking=True)
        return loss
    return step, step_e2e, host_x.numel() * 4, 4


def build_cls_oracle(size, batch, seed):
    from cls_case import build_cls_model, oracle_cls_logits
    from helpers import synth_images
    _, sd = build_cls_model("global", 21, size=size)
    leaf = {k: v.clone().requires_grad_(not k.endswith("pos_emb")) for k, v in sd.items()}
    x = synth_images(batch, ["bscan"], seed=1234)["bscan"]
    tgt = torch.arange(batch) % 5
    depth_heads = (12, 12) if size == "base" else (24, 16)

    def run():
        from oracle import mirage_oracle as O
        msd = {k[len("model."):]: v for k, v in leaf.items() if k.startswith("model.")}
        tok = O.light_forward({"bscan": x}, msd, *depth_heads)
        loss = torch.nn.functional.cross_entropy(O.cls_head(tok, leaf, "global"), tgt)
        loss.backward()
        for v in leaf.values():
            v.grad = None
        return loss
    return run


def build_pretrain_oracle(size, mods, batch, seed):
    from helpers import synth_images, synth_state_dict
    from pretrain_case import build_pretrain_model, oracle_step, sample_masks
    model, _ = build_pretrain_model(size, mods)
    sd = model.state_dict()
    sd.update(synth_state_dict({k: v.shape for k, v in sd.items()}, 3))
    x = synth_images(batch, mods, seed=1234)
    masks = sample_masks(model, batch, 98, seed=100)

    def run():
        return oracle_step(sd, x, masks, size, mods)
    return run
