"""Optimizer step, gradient norm and schedules around the MIRAGE hot path.

Same names and semantics as the reference's ``mutils/optim_factory.py`` (parameter groups, layer-wise lr
decay, ``create_optimizer``), ``mutils/native_scaler.py`` (``NativeScalerWithGradNormCount``,
``get_grad_norm_``, ``cosine_scheduler``) -- the arithmetic runs in ``csrc/optim.cu``: ONE pass over the
parameters does the AdamW update in fp32, rewrites the bf16 weight shadows the GEMMs read, zeroes the
gradient for the next step and produces the global gradient norm; clipping / skipping need one extra
read of the gradient buckets.  All scalars a step depends on (step count, bias corrections, clip
coefficient, lr / wd per group) live in device memory, so the whole training step -- optimizer included
-- can sit inside one CUDA graph while the host keeps driving the schedule.

bf16 tensor-core operands with fp32 accumulation and an fp32 residual stream need no loss scaling, so the
reference's ``GradScaler`` (fp16 autocast) has no counterpart: ``loss_scale`` is reported as 1.0.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, Optional

import numpy as np
import torch

from . import _lib as L
from . import functional as Fn
from . import ops


# ---------------------------------------------------------------------------------------------
# parameter groups (mutils/optim_factory.py:6-92)
# ---------------------------------------------------------------------------------------------
def get_num_layer_for_vit(var_name: str, num_max_layer: int) -> int:
    """Layer id used for layer-wise lr decay (optim_factory.py:6-20)."""
    if var_name in ("cls_token", "mask_token", "pos_embed", "global_tokens"):
        return 0
    if var_name.startswith("patch_embed") or var_name.startswith("input_adapters"):
        return 0
    if var_name.startswith("rel_pos_bias"):
        return num_max_layer - 1
    if var_name.startswith("blocks") or var_name.startswith("encoder"):
        return int(var_name.split('.')[1]) + 1
    return num_max_layer - 1


class LayerDecayValueAssigner:
    def __init__(self, values):
        self.values = values

    def get_scale(self, layer_id):
        return self.values[layer_id]

    def get_layer_id(self, var_name):
        return get_num_layer_for_vit(var_name, len(self.values))


def get_parameter_groups(model, weight_decay=1e-5, skip_list=(), get_num_layer=None, get_layer_scale=None,
                         decoder_decay=None, decoder_list=(), no_lr_scale_list=()):
    """Groups keyed by (layer id,) decay class; 1-D tensors, ``.bias`` and ``skip_list`` names get no decay
    (optim_factory.py:33-92).  Returns a list of ``{'params', 'weight_decay', 'lr_scale'}`` dicts."""
    groups: dict = {}
    for name, param in model.named_parameters():
        if not param.requires_grad:
            continue
        if param.ndim == 1 or name.endswith(".bias") or name in skip_list:
            key, wd = "no_decay", 0.
        elif decoder_decay is not None and (name.startswith("decoder.") or name in decoder_list):
            key, wd = "decoder_decay", decoder_decay
        else:
            key, wd = "decay", weight_decay
        layer_id, skip_scale = None, False
        if get_num_layer is not None:
            layer_id = get_num_layer(name)
            key = "layer_%d_%s" % (layer_id, key)
            if name in no_lr_scale_list:
                skip_scale = True
                key = f"{key}_no_lr_scale"
        if key not in groups:
            scale = get_layer_scale(layer_id) if (get_layer_scale is not None and not skip_scale) else 1.
            groups[key] = {"weight_decay": wd, "params": [], "lr_scale": scale, "names": []}
        groups[key]["params"].append(param)
        groups[key]["names"].append(name)
    out = []
    for g in groups.values():
        g.pop("names")
        out.append(g)
    return out


def cosine_scheduler(base_value, final_value, epochs, niter_per_ep, warmup_epochs=0, start_warmup_value=0,
                     warmup_steps=-1):
    """Per-iteration values: linear warm-up then half-cosine (native_scaler.py:64-88)."""
    warmup_iters = warmup_epochs * niter_per_ep
    if warmup_steps > 0:
        warmup_iters = warmup_steps
    warm = np.linspace(start_warmup_value, base_value, warmup_iters) if warmup_epochs > 0 else np.array([])
    n = epochs * niter_per_ep - warmup_iters
    i = np.arange(n)
    main = np.array([final_value + 0.5 * (base_value - final_value) * (1 + math.cos(math.pi * k / n)) for k in i])
    sched = np.concatenate((warm, main))
    assert len(sched) == epochs * niter_per_ep
    return sched


def assign_step_hyper(optimizer, it: int, lr_schedule_values=None, wd_schedule_values=None):
    """The per-step assignment of run_pretraining.py:683-688."""
    for group in optimizer.param_groups:
        if lr_schedule_values is not None:
            group['lr'] = lr_schedule_values[it] * group.get('lr_scale', 1.0)
        if wd_schedule_values is not None and group['weight_decay'] > 0:
            group['weight_decay'] = wd_schedule_values[it]


def get_grad_norm_(parameters, norm_type: float = 2.0) -> torch.Tensor:
    """Global gradient norm (native_scaler.py:46-61); plain tensor ops -- FusedAdamW.step() produces the same
    number as a by-product of its update pass."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad.detach() for p in parameters if p.grad is not None]
    if not grads:
        return torch.tensor(0.)
    if norm_type == math.inf:
        return max(g.abs().max() for g in grads)
    return torch.norm(torch.stack([torch.norm(g, norm_type) for g in grads]), norm_type)


# ---------------------------------------------------------------------------------------------
# fused AdamW
# ---------------------------------------------------------------------------------------------
class FusedAdamW(torch.optim.Optimizer):
    """``torch.optim.AdamW`` arithmetic (fp32 state, decoupled decay) in one kernel pass.

    ``param_groups`` behave as in torch (``lr``, ``weight_decay``, ``betas``, ``eps``; extra keys such as
    ``lr_scale`` are kept for the caller's schedule loop).  ``betas`` / ``eps`` must be common to all groups.
    ``step(clip_grad=None, skip_grad=None)`` returns the global L2 norm of the gradients as a 0-d CUDA tensor
    (no host sync).  ``zero_grad_in_step=True`` additionally zeroes every gradient inside the same pass (for
    trainers whose gradients live in persistent buffers, e.g. ``ddp.GradBucketAllReduce``).
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2,
                 zero_grad_in_step: bool = False, grad_buckets: Optional[Iterable[torch.Tensor]] = None):
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.zero_grad_in_step = zero_grad_in_step
        self.grad_buckets = list(grad_buckets) if grad_buckets is not None else None
        self._table = None
        self._dev = None
        self._pending_step = None      # step count to push to the device (after load_state_dict)

    # -- tables -----------------------------------------------------------------------------------
    def _live_params(self):
        for gi, group in enumerate(self.param_groups):
            for p in group['params']:
                if p.requires_grad and p.grad is not None:
                    yield gi, p

    def _signature(self):
        return (Fn.shadow_epoch(), tuple((id(p), p.data_ptr(), p.grad.data_ptr()) for _, p in self._live_params()))

    def _build(self):
        live = list(self._live_params())
        if not live:
            raise L.MirageB200Error("FusedAdamW.step(): no parameter has a gradient")
        dev = live[0][1].device
        if dev.type != 'cuda':
            raise L.MirageB200Error("FusedAdamW needs CUDA parameters (no CPU fallback)")
        segs = (L.OptimSegment * len(live))()
        prefix = np.zeros(len(live) + 1, dtype=np.int32)
        for i, (gi, p) in enumerate(live):
            if p.dtype != torch.float32 or not p.is_contiguous() or p.grad.dtype != torch.float32 \
                    or not p.grad.is_contiguous():
                raise L.MirageB200Error("FusedAdamW: parameters and gradients must be contiguous fp32")
            st = self.state[p]
            if 'exp_avg' not in st:
                st['step'] = torch.zeros((), dtype=torch.float32, device=dev)
                st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            sh = Fn.shadow_of(p)
            s = segs[i]
            s.param, s.grad = p.data_ptr(), p.grad.data_ptr()
            s.exp_avg, s.exp_avg_sq = st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr()
            s.shadow = sh.data_ptr() if sh is not None else None
            s.numel = p.numel()
            s.group = gi
            ptrs = [s.param, s.grad, s.exp_avg, s.exp_avg_sq]
            aligned = all(q % 16 == 0 for q in ptrs) and (sh is None or sh.data_ptr() % 8 == 0)
            s.flags = 1 if aligned else 0
            prefix[i + 1] = prefix[i] + L.lib().mb_optim_blocks(p.numel())
        raw = bytes(segs)
        seg_dev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        prefix_dev = torch.from_numpy(prefix).to(dev)
        n_blocks = int(prefix[-1])
        n_groups = len(self.param_groups)
        tab = {
            'sig': self._signature(), 'n_segs': len(live), 'n_blocks': n_blocks, 'segs': seg_dev,
            'prefix': prefix_dev,
            'hyper_host': torch.zeros((n_groups, 4), dtype=torch.float32).pin_memory(),
            'hyper_dev': torch.zeros((n_groups, 4), dtype=torch.float32, device=dev),
            'partials': torch.zeros(max(n_blocks, 4096), dtype=torch.float32, device=dev),
            'numel': sum(p.numel() for _, p in live),
        }
        if self._table is None:
            # {step i64, bias_corr1, bias_corr2_sqrt, clip_coef, grad_norm, skipped i32, pad} = 32 bytes
            self._state_dev = torch.zeros(8, dtype=torch.int32, device=dev)
        self._table = tab
        self._dev = dev

    def _push_hyper(self):
        tab = self._table
        h = tab['hyper_host']
        for gi, g in enumerate(self.param_groups):
            h[gi, 0] = float(g['lr'])
            h[gi, 1] = 1.0    # lr_scale is folded into group['lr'] by the caller's loop (run_pretraining.py:686)
            h[gi, 2] = float(g['weight_decay'])
        tab['hyper_dev'].copy_(h, non_blocking=True)

    def refresh_hyper(self):
        """Copy the groups' current lr / weight_decay to the device table.  ``step()`` does it itself; call this
        before replaying a CUDA graph that contains ``step()``."""
        if self._table is None:
            self._build()
        self._push_hyper()

    # -- step ---------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None, clip_grad: Optional[float] = None, skip_grad: Optional[float] = None):
        assert closure is None, "FusedAdamW does not take a closure"
        if self._table is None or (not torch.cuda.is_current_stream_capturing()
                                   and self._table['sig'] != self._signature()):
            self._build()
        tab = self._table
        g0 = self.param_groups[0]
        beta1, beta2 = g0['betas']
        eps = g0['eps']
        for g in self.param_groups:
            if tuple(g['betas']) != (beta1, beta2) or g['eps'] != eps:
                raise L.MirageB200Error("FusedAdamW: betas / eps must be the same in every group")
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:
            self._push_hyper()
            if self._pending_step is not None:
                self._state_dev[:2].copy_(torch.tensor([self._pending_step], dtype=torch.int64).view(torch.int32))
                self._pending_step = None
        stream = torch.cuda.current_stream().cuda_stream
        lib = L.lib()
        st_ptr = self._state_dev.data_ptr()
        clip = float(clip_grad) if clip_grad is not None else 0.0
        skip = float(skip_grad) if (skip_grad is not None and clip_grad is None) else 0.0
        two_pass = clip > 0.0 or skip > 0.0
        with ops._rec("adamw_step", 34.0 * tab['numel'], "byte", kernels=3 if not two_pass else 2 + 1):
            if two_pass:
                n_part = self._sumsq_partials(tab, stream)
                L.check(lib.mb_optim_prepare(st_ptr, tab['partials'].data_ptr(), n_part, clip, skip, beta1, beta2,
                                             stream), "mb_optim_prepare")
                L.check(lib.mb_adamw_step(tab['segs'].data_ptr(), tab['prefix'].data_ptr(), tab['n_segs'],
                                          tab['n_blocks'], tab['hyper_dev'].data_ptr(), st_ptr, beta1, beta2, eps,
                                          1 if self.zero_grad_in_step else 0, None, stream), "mb_adamw_step")
            else:
                L.check(lib.mb_optim_prepare(st_ptr, None, 0, 0.0, 0.0, beta1, beta2, stream), "mb_optim_prepare")
                L.check(lib.mb_adamw_step(tab['segs'].data_ptr(), tab['prefix'].data_ptr(), tab['n_segs'],
                                          tab['n_blocks'], tab['hyper_dev'].data_ptr(), st_ptr, beta1, beta2, eps,
                                          1 if self.zero_grad_in_step else 0, tab['partials'].data_ptr(), stream),
                        "mb_adamw_step")
                L.check(lib.mb_optim_finish(st_ptr, tab['partials'].data_ptr(), tab['n_blocks'], stream),
                        "mb_optim_finish")
        return self._state_dev.view(torch.float32)[5].clone()

    def _sumsq_partials(self, tab, stream) -> int:
        """Sum of squares of every gradient into consecutive partials; returns how many were written."""
        lib = L.lib()
        part = tab['partials']
        if self.grad_buckets is not None:
            bufs = self.grad_buckets
        else:
            bufs = [p.grad for _, p in self._live_params()]
        per = max(1, min(512, part.numel() // max(1, len(bufs))))
        off = 0
        for b in bufs:
            n = max(1, min(per, (b.numel() + 4095) // 4096))
            if b.data_ptr() % 16 != 0:
                raise L.MirageB200Error("FusedAdamW: gradient buffers must be 16-byte aligned for clipping")
            L.check(lib.mb_sumsq(b.data_ptr(), b.numel(), part.data_ptr() + 4 * off, n, stream), "mb_sumsq")
            ops._Stats.launches += 1
            off += n
        if off > part.numel():
            raise L.MirageB200Error("FusedAdamW: partial buffer too small")
        return off

    # -- introspection ------------------------------------------------------------------------------
    @property
    def grad_norm(self) -> torch.Tensor:
        return self._state_dev.view(torch.float32)[5]

    @property
    def step_count(self) -> torch.Tensor:
        """Completed (non-skipped) steps, 0-d int64 CUDA tensor."""
        return self._state_dev[:2].view(torch.int64)[0]

    @property
    def last_step_skipped(self) -> torch.Tensor:
        return self._state_dev[6]

    # -- checkpoint compatibility with torch.optim.AdamW ------------------------------------------------
    def state_dict(self):
        if self._table is not None:
            n = float(self.step_count.item())
            for st in self.state.values():
                if 'step' in st:
                    st['step'].fill_(n)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        steps = [float(st['step']) for st in self.state.values() if 'step' in st]
        self._pending_step = int(max(steps)) if steps else 0
        if self._table is not None:
            self._table['sig'] = None      # the state tensors were replaced: rebuild the pointer table


def create_optimizer(args, model, get_num_layer=None, get_layer_scale=None, filter_bias_and_bn=True,
                     skip_list=None, grad_buckets=None, zero_grad_in_step=False):
    """mutils/optim_factory.py:95-209 for the optimizers the MIRAGE recipes use (``adamw``: the fused kernel;
    ``sgd`` / ``momentum`` / ``adam`` fall through to torch -- they are not on the benchmarked path)."""
    opt_lower = args.opt.lower()
    weight_decay = args.weight_decay
    decoder_decay = getattr(args, 'decoder_decay', None)
    try:
        no_lr_scale_list = args.no_lr_scale_list.split('-')
    except AttributeError:
        no_lr_scale_list = []
    if weight_decay and filter_bias_and_bn:
        skip = skip_list if skip_list is not None else (model.no_weight_decay() if hasattr(model, 'no_weight_decay') else {})
        decoder = model.decoder_weight_decay() if hasattr(model, 'decoder_weight_decay') else {}
        parameters = get_parameter_groups(model, weight_decay, skip, get_num_layer, get_layer_scale, decoder_decay,
                                          decoder, no_lr_scale_list)
        weight_decay = 0.
    else:
        parameters = [p for p in model.parameters() if p.requires_grad]
    opt_args = dict(lr=args.lr, weight_decay=weight_decay)
    if getattr(args, 'opt_eps', None) is not None:
        opt_args['eps'] = args.opt_eps
    if getattr(args, 'opt_betas', None) is not None:
        opt_args['betas'] = tuple(args.opt_betas)
    opt_lower = opt_lower.split('_')[-1]
    if opt_lower == 'adamw':
        return FusedAdamW(parameters, grad_buckets=grad_buckets, zero_grad_in_step=zero_grad_in_step, **opt_args)
    if opt_lower in ('sgd', 'nesterov'):
        opt_args.pop('eps', None)
        return torch.optim.SGD(parameters, momentum=args.momentum, nesterov=True, **opt_args)
    if opt_lower == 'momentum':
        opt_args.pop('eps', None)
        return torch.optim.SGD(parameters, momentum=args.momentum, nesterov=False, **opt_args)
    if opt_lower == 'adam':
        return torch.optim.Adam(parameters, **opt_args)
    raise ValueError(f"Unknown optimizer {args.opt}")


class NativeScalerWithGradNormCount:
    """Call-compatible stand-in for mutils/native_scaler.py:10-44: backward, gradient exchange (when a
    ``GradBucketAllReduce`` is attached), norm / clip / skip and the optimizer step.  There is no loss scale in
    bf16; ``state_dict()`` reports ``scale = 1.0`` so the logging code of run_pretraining.py:743 keeps working."""
    state_dict_key = "amp_scaler"

    def __init__(self, enabled=True, ddp=None):
        self.ddp = ddp

    def __call__(self, loss, optimizer, clip_grad=None, skip_grad=None, parameters=None, create_graph=False,
                 update_grad=True):
        loss.backward(create_graph=create_graph)
        if not update_grad:
            return None
        if self.ddp is not None:
            self.ddp.finish()
        if isinstance(optimizer, FusedAdamW):
            return optimizer.step(clip_grad=clip_grad, skip_grad=skip_grad)
        if clip_grad is not None:
            assert parameters is not None
            norm = torch.nn.utils.clip_grad_norm_(parameters, clip_grad)
        else:
            norm = get_grad_norm_(parameters)
            if skip_grad is not None and norm >= skip_grad:
                return norm
        optimizer.step()
        return norm

    def state_dict(self):
        return {"scale": 1.0}

    def load_state_dict(self, state_dict):
        pass
