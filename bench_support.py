"""Training workloads for bench.py (cfg 3 / 4 / 5 of BASELINE.json): model + criteria + optimizer step +
gradient all-reduce on the GPU arm, and the oracle's forward+backward on the CPU arm (which imports
nothing from the product package)."""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "tests"))


def _make_optimizer(model, ddp, lr, opts, skip=()):
    """AdamW over the reference's parameter groups (mutils/optim_factory.py:33-92): 1-D tensors, biases and the
    model's no_weight_decay() names undecayed; beta = (0.9, 0.95), wd 0.05 (run_pretraining.py:172,189)."""
    from mirage_b200.optim import FusedAdamW, get_parameter_groups
    groups = get_parameter_groups(model, 0.05, skip)
    if opts.optimizer == "fused":
        return FusedAdamW(groups, lr=lr, betas=(0.9, 0.95), zero_grad_in_step=True,
                          grad_buckets=[b.flat for b in ddp.buckets])
    return torch.optim.AdamW(groups, lr=lr, betas=(0.9, 0.95), fused=True, capturable=True)


def _schedule(lr, steps=4096):
    from mirage_b200.optim import cosine_scheduler
    return cosine_scheduler(lr, 0.0, 1, steps, warmup_epochs=0), cosine_scheduler(0.05, 0.05, 1, steps)


def _train_step_factory(model, ddp, opt, opts, loss_fn):
    """Returns one_step(x): zero / forward / loss / backward (+ bucketed all-reduce) / optimizer step."""
    fused = opts.optimizer == "fused"

    def one_step(x):
        ddp.zero_grad(memset=not fused)       # FusedAdamW zeroes the buckets inside its own update pass
        loss = loss_fn(x)
        loss.backward()
        ddp.finish()
        opt.step()
        return loss
    return one_step


def _host_schedule(opt, opts, lr):
    """The per-step lr / weight-decay assignment of run_pretraining.py:683-688 (host side, every step)."""
    from mirage_b200.optim import assign_step_hyper
    lr_sched, wd_sched = _schedule(lr)
    state = {"it": 0}

    def tick():
        it = state["it"] % len(lr_sched)
        state["it"] += 1
        assign_step_hyper(opt, it, lr_sched, wd_sched)
        if opts.optimizer == "fused":
            opt.refresh_hyper()
    return tick


def build_pretrain_step(size, mods, per_gpu, dev, rank, world, opts):
    from helpers import load_synth, synth_images
    from mirage_b200.ddp import GradBucketAllReduce
    from pretrain_case import build_criteria, build_pretrain_model

    model, _ = build_pretrain_model(size, mods)
    load_synth(model, seed=3)
    model = model.to(dev).train()
    if opts.mask_sampler == "device":
        model.set_mask_sampler("device", seed=4242 + rank)
    crits = build_criteria(mods)
    ddp = GradBucketAllReduce(model, bucket_mb=opts.bucket_mb, reserve_sms=opts.reserve_sms if world > 1 else 0)
    lr = 1e-4 * per_gpu * world / 256
    opt = _make_optimizer(model, ddp, lr, opts, model.no_weight_decay())
    tick = _host_schedule(opt, opts, lr)

    base = synth_images(8, mods, seed=1234 + rank)
    reps = per_gpu // 8 + 1
    host_in = {k: v.repeat(reps, *([1] * (v.dim() - 1)))[:per_gpu].contiguous().pin_memory() for k, v in base.items()}
    dev_in = {k: v.to(dev) for k, v in host_in.items()}
    torch.manual_seed(100 + rank)
    host_loss = torch.zeros(1).pin_memory()

    def loss_fn(x):
        preds, masks = model(x, num_encoded_tokens=98, alphas=1.0, sample_tasks_uniformly=False)
        return sum(crits[d](preds[d].float(), x[d], mask=masks[d]) for d in mods)
    one_step = _train_step_factory(model, ddp, opt, opts, loss_fn)

    def step():
        tick()
        return one_step(dev_in)

    def step_e2e():
        tick()
        x = {k: v.to(dev, non_blocking=True) for k, v in host_in.items()}
        loss = one_step(x)
        host_loss.copy_(loss.detach().reshape(1), non_blocking=True)
        return loss

    def graph_hooks():
        """Whole-step CUDA graph: forward + losses + backward (+ the NCCL all-reduce of every bucket, which
        torch captures as cross-stream dependencies) + the optimizer step are captured once.  With the
        on-device mask sampler the masks are drawn INSIDE the graph (the kernel advances its own Philox
        counter); with the reference sampler they are drawn eagerly before every replay exactly as the
        reference does (CPU Dirichlet + device noise, mirage/model.py:168-239) and copied into the fixed
        tensors the graph reads."""
        from mirage_b200.graphs import GraphedCallable
        x_fixed = {k: v.clone() for k, v in dev_in.items()}
        resample = None
        if opts.mask_sampler != "device":
            sampler = model.generate_random_masks
            toks = {d: torch.empty(per_gpu, 256, 0, device=dev) for d in mods}
            tm0, keep0, restore0 = sampler(toks, 98, alphas=1.0)
            fixed = ({d: t.clone() for d, t in tm0.items()}, keep0.clone(), restore0.clone())

            def resample():
                tm, keep, restore = sampler(toks, 98, alphas=1.0)
                for d in mods:
                    fixed[0][d].copy_(tm[d])
                fixed[1].copy_(keep)
                fixed[2].copy_(restore)
            model.generate_random_masks = lambda *a, **k: fixed
        g = GraphedCallable(lambda: one_step(x_fixed), refresh_weights=opts.optimizer != "fused").capture()

        def gstep():
            tick()
            if resample is not None:
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
            tick()
            if resample is not None:
                resample()
            loss = g()
            host_loss.copy_(loss.detach().reshape(1), non_blocking=True)
            return loss
        return gstep, gstep_e2e

    h2d = sum(v.numel() * v.element_size() for v in host_in.values())
    return step, step_e2e, h2d, 4, graph_hooks, ddp


def build_cls_step(size, per_gpu, dev, rank, world, opts):
    """cfg 5: miragecls_factory['global'] fine-tune step (forward + CE + backward + gradient exchange + AdamW),
    bscan 512x512, 5 classes.  ``opts.linear_probe``: only ``head.*`` trains
    (fm_cls_config.py:111-124)."""
    from cls_case import build_cls_model
    from helpers import synth_images
    from mirage_b200.ddp import GradBucketAllReduce
    model, _ = build_cls_model("global", 21, device=dev, size=size)
    model.train()
    if opts.mask_sampler == "device":
        model.model.set_mask_sampler("device", seed=777 + rank)
    if getattr(opts, "linear_probe", False):
        for n, p in model.named_parameters():
            p.requires_grad_(n.startswith("head."))
    ddp = GradBucketAllReduce(model, bucket_mb=opts.bucket_mb, reserve_sms=opts.reserve_sms if world > 1 else 0)
    opt = _make_optimizer(model, ddp, 1e-4, opts)
    tick = _host_schedule(opt, opts, 1e-4)
    base = synth_images(8, ["bscan"], seed=1234 + rank)["bscan"]
    host_x = base.repeat(per_gpu // 8 + 1, 1, 1, 1)[:per_gpu].contiguous().pin_memory()
    dev_x = host_x.to(dev)
    tgt = torch.randint(0, 5, (per_gpu,), generator=torch.Generator().manual_seed(5 + rank)).to(dev)
    torch.manual_seed(100 + rank)
    host_loss = torch.zeros(1).pin_memory()
    smoothing = float(getattr(opts, "label_smoothing", 0.0))

    def loss_fn(x):
        return torch.nn.functional.cross_entropy(model(x).float(), tgt, label_smoothing=smoothing)
    one_step = _train_step_factory(model, ddp, opt, opts, loss_fn)

    def step():
        tick()
        return one_step(dev_x)

    def step_e2e():
        tick()
        loss = one_step(host_x.to(dev, non_blocking=True))
        host_loss.copy_(loss.detach().reshape(1), non_blocking=True)
        return loss

    def graph_hooks():
        """Whole-step CUDA graph (needs the on-device mask sampler: the cls wrappers draw a random token
        permutation every forward)."""
        if opts.mask_sampler != "device":
            raise RuntimeError("cls step capture needs --mask-sampler device")
        from mirage_b200.graphs import GraphedCallable
        x_fixed = dev_x.clone()
        g = GraphedCallable(lambda: one_step(x_fixed), refresh_weights=opts.optimizer != "fused").capture()

        def gstep():
            tick()
            return g()

        def gstep_e2e():
            tick()
            x_fixed.copy_(host_x, non_blocking=True)
            loss = g()
            host_loss.copy_(loss.detach().reshape(1), non_blocking=True)
            return loss
        return gstep, gstep_e2e
    return step, step_e2e, host_x.numel() * 4, 4, graph_hooks, ddp


# ---------------------------------------------------------------------------------------------
# CPU legs (oracle only: no product import)
# ---------------------------------------------------------------------------------------------
def build_cls_oracle(size, batch, seed):
    from helpers import synth_images, synth_state_dict
    from oracle import mirage_oracle as O
    sd = O.cls_state_dict_shapes(size)
    sd.update(synth_state_dict({k: v.shape for k, v in sd.items()}, 21))
    leaf = {k: v.clone().requires_grad_(not k.endswith("pos_emb")) for k, v in sd.items()}
    x = synth_images(batch, ["bscan"], seed=1234)["bscan"]
    tgt = torch.arange(batch) % 5
    _, depth, heads = O.MODEL_SIZES[size]

    def run():
        msd = {k[len("model."):]: v for k, v in leaf.items() if k.startswith("model.")}
        tok = O.light_forward({"bscan": x}, msd, depth, heads)
        loss = torch.nn.functional.cross_entropy(O.cls_head(tok, leaf, "global"), tgt)
        loss.backward()
        for v in leaf.values():
            v.grad = None
        return loss
    return run


def build_pretrain_oracle(size, mods, batch, seed):
    from helpers import synth_images, synth_state_dict
    from oracle import mirage_oracle as O
    sd = O.pretrain_state_dict_shapes(mods, size)
    sd.update(synth_state_dict({k: v.shape for k, v in sd.items()}, 3))
    x = synth_images(batch, mods, seed=1234)
    torch.manual_seed(100)
    tm, keep, restore = O.random_masks([256] * len(mods), batch, 98, alphas=1.0)
    task_masks = dict(zip(mods, tm))
    _, depth, heads = O.MODEL_SIZES[size]

    def run():
        leaf = {k: v.clone().requires_grad_(not k.endswith("pos_emb")) for k, v in sd.items()}
        preds, _ = O.pretrain_forward(x, leaf, depth, heads, (None, keep, restore), mods)
        total, _losses = O.pretrain_loss(preds, x, task_masks)
        total.backward()
        return total
    return run
